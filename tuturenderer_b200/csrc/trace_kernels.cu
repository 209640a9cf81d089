// trace_kernels.cu — the kernels that walk the compressed wide tree (wide.cuh): ray batches, the wavefront's
// extend / shadow queues and the BDPT queues.  One ray per lane, packets pulled from the queue by a warp-level
// cursor (persistent grid = SMs x resident blocks), traversal stack in shared memory ((wide_depth + 1) 64-bit
// words per thread).  Rays the wide walk does not take (an infinite 1/d, scenes without a wide tree) fall back to
// the binary walk inside trace_ray().  sm_100a only.
#include "trace_kernels.hpp"

#include "wide.cuh"

namespace tutu {
namespace {

__device__ __forceinline__ bool next_packet(unsigned long long* cursor, unsigned long long n, unsigned long long step,
                                            unsigned long long& base) {
  unsigned long long b = 0;
  if ((threadIdx.x & 31u) == 0) b = atomicAdd(cursor, step);
  base = __shfl_sync(0xFFFFFFFFu, b, 0);
  return base < n;
}

template <bool ANY>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_wide(const __grid_constant__ DevScene sc, const float4* __restrict__ rays, unsigned long long n, TutuHit* __restrict__ out,
             uint8_t* __restrict__ out_any, unsigned long long* __restrict__ next, const unsigned* __restrict__ perm) {
  extern __shared__ unsigned long long s_stack[];
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base;
  while (next_packet(next, n, 32ull, base)) {
    const unsigned long long j = base + lane;
    if (j < n) {
      const unsigned long long i = perm ? (unsigned long long)__ldg(perm + j) : j;
      const float4 o = __ldg(rays + 2 * i);
      const float4 d = __ldg(rays + 2 * i + 1);
      Hit h;
      const bool any = trace_ray<ANY, 0>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, ANY ? d.w : 0.f, h, s_stack + threadIdx.x, blockDim.x);
      if (ANY) {
        out_any[i] = any ? 1 : 0;
      } else {
        const int prim = h.slot >= 0 ? __ldg(sc.slot_to_prim + (h.slot & (int)kSlotMask)) : -1;
        reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), h.t, h.u, h.v);
      }
    }
    __syncwarp();
  }
}

template <bool ANY>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_count_wide(const __grid_constant__ DevScene sc, const float4* __restrict__ rays, unsigned long long n,
                   unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned long long s_stack[];
  unsigned long long nodes = 0, prims = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float4 o = __ldg(rays + 2 * i);
    const float4 d = __ldg(rays + 2 * i + 1);
    const Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
    const RayPre p = make_pre(r);
    Hit h;
    h.t = FLT_MAX, h.u = 0.f, h.v = 0.f, h.slot = -1;
    VisitCount vc;
    if (ray_is_regular(p)) {  // irregular rays walk the reference's binary tree and are not counted here
      float te;
      if (box_test_regular(p, sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2], te))
        walk_wide<ANY, true>(sc, r, p, ANY ? d.w : 0.f, h, s_stack + threadIdx.x, blockDim.x, &vc);
    }
    nodes += vc.nodes;
    prims += vc.prims;
  }
  for (int o = 16; o > 0; o >>= 1) {
    nodes += __shfl_down_sync(0xFFFFFFFFu, nodes, o);
    prims += __shfl_down_sync(0xFFFFFFFFu, prims, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counts + 0, nodes);
    atomicAdd(counts + 1, prims);
  }
}

// IIntersectStrategy::UpdateInter -> getIntersection for the wavefront's ray queue
__global__ void __launch_bounds__(kTraceBlock)
wf_extend_wide(const __grid_constant__ DevScene sc, WfBuffers b, int cur) {
  extern __shared__ unsigned long long s_stack[];
  const unsigned n = b.ctl->n_cur;
  const float4* __restrict__ ro = b.ray_o[cur];
  const float4* __restrict__ rd = b.ray_d[cur];
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base;
  while (next_packet(&b.ctl->cursor_extend, n, kPacketRays, base)) {
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned i = (unsigned)base + k + lane;
      if (i < n) {
        const float4 o = __ldcs(ro + i);
        const float4 d = __ldcs(rd + i);
        Hit h;
        if (__float_as_uint(o.w) == kDeadQueueEntry)  // unused end of a reserved chunk (wf_types.cuh)
          h.t = 0.f, h.u = 0.f, h.v = 0.f, h.slot = kDeadSlot;
        else
          trace_ray<false, 1>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h, s_stack + threadIdx.x, blockDim.x);
        __stcs(b.hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
      }
      __syncwarp();
    }
  }
}

// isShadowRayBlocked -> hasIntersection, then the deferred NEE add (same record handling as wavefront.cuh: wf_shadow)
__global__ void __launch_bounds__(kTraceBlock)
wf_shadow_wide(const __grid_constant__ DevScene sc, WfBuffers b, int nxt) {
  extern __shared__ unsigned long long s_stack[];
  const unsigned n = b.ctl->n_shadow;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base;
  while (next_packet(&b.ctl->cursor_shadow, n, kPacketRays, base)) {
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned j = (unsigned)base + k + lane;
      float4 o = make_float4(0.f, 0.f, 0.f, __uint_as_float(kDeadQueueEntry)), d = o;
      if (j < n) o = __ldcs(b.sh_o + j), d = __ldcs(b.sh_d + j);
      if (__float_as_uint(o.w) != kDeadQueueEntry) {
        Hit h;
        const bool blocked = trace_ray<true, 2>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h, s_stack + threadIdx.x, blockDim.x);
        const unsigned dst = __float_as_uint(d.w);
        if (dst == kShadowFinalDst) {
          const float4 c = __ldcs(b.sh_c + j);
          const float4 L4 = __ldcs(b.sh_L + j);
          f3 L = mk(L4.x, L4.y, L4.z);
          if (!blocked) L = L + mk(c.x, c.y, c.z);
          accum_add(b.accum, b.ctl, __float_as_uint(c.w), L);
        } else if (!blocked) {
          const float4 c = __ldcs(b.sh_c + j);
          float4 s = b.st2[nxt][dst];
          s.x += c.x, s.y += c.y, s.z += c.z;
          b.st2[nxt][dst] = s;
        }
      }
      __syncwarp();
    }
  }
}

// BDPT walk queue (bdpt.cuh): closest hits of both sub-path kinds
__global__ void __launch_bounds__(kTraceBlock)
q_extend_wide(const __grid_constant__ DevScene sc, const float4* __restrict__ ro, const float4* __restrict__ rd,
              float4* __restrict__ hit, const unsigned* __restrict__ n_ptr, unsigned long long* cursor) {
  extern __shared__ unsigned long long s_stack[];
  const unsigned n = *n_ptr;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base;
  while (next_packet(cursor, n, kPacketRays, base)) {
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned i = (unsigned)base + k + lane;
      if (i < n) {
        const float4 o = __ldcs(ro + i);
        const float4 d = __ldcs(rd + i);
        Hit h;
        h.t = FLT_MAX, h.u = 0.f, h.v = 0.f, h.slot = -1;
        if (__float_as_uint(d.w) != kDeadQueueEntry)
          trace_ray<false, 3>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h, s_stack + threadIdx.x, blockDim.x);
        __stcs(hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
      }
      __syncwarp();
    }
  }
}

// BDPT connection rays: unoccluded connections add their weighted contribution to the frame buffer
__global__ void __launch_bounds__(kTraceBlock)
q_shadow_add_wide(const __grid_constant__ DevScene sc, const float4* __restrict__ so, const float4* __restrict__ sd,
                  const float4* __restrict__ scn, float* __restrict__ accum, const unsigned* __restrict__ n_ptr,
                  unsigned long long* cursor) {
  extern __shared__ unsigned long long s_stack[];
  const unsigned n = *n_ptr;
  const unsigned lane = threadIdx.x & 31u;
  unsigned long long base;
  while (next_packet(cursor, n, kPacketRays, base)) {
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned j = (unsigned)base + k + lane;
      if (j < n) {
        const float4 o = __ldcs(so + j);
        const float4 d = __ldcs(sd + j);
        Hit h;
        if (!trace_ray<true, 4>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h, s_stack + threadIdx.x, blockDim.x)) {
          const float4 c = __ldcs(scn + j);
          float* p = accum + (size_t)__float_as_uint(d.w) * 3;
          atomicAdd(p + 0, c.x);
          atomicAdd(p + 1, c.y);
          atomicAdd(p + 2, c.z);
        }
      }
      __syncwarp();
    }
  }
}

template <class K>
cudaError_t grid_of(K kernel, int sm_count, size_t smem, int* grid) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTraceBlock, smem);
  *grid = sm_count * (per_sm < 1 ? 1 : per_sm);
  return e;
}

}  // namespace

cudaError_t wide_grids(int sm_count, size_t smem, WideGrids* g) {
  cudaError_t e;
  if ((e = grid_of(k_trace_wide<false>, sm_count, smem, &g->batch_closest)) != cudaSuccess) return e;
  if ((e = grid_of(k_trace_wide<true>, sm_count, smem, &g->batch_any)) != cudaSuccess) return e;
  if ((e = grid_of(wf_extend_wide, sm_count, smem, &g->wf_extend)) != cudaSuccess) return e;
  if ((e = grid_of(wf_shadow_wide, sm_count, smem, &g->wf_shadow)) != cudaSuccess) return e;
  if ((e = grid_of(q_extend_wide, sm_count, smem, &g->q_extend)) != cudaSuccess) return e;
  if ((e = grid_of(q_shadow_add_wide, sm_count, smem, &g->q_shadow)) != cudaSuccess) return e;
  int unused;
  if ((e = grid_of(k_trace_count_wide<false>, sm_count, smem, &unused)) != cudaSuccess) return e;
  if ((e = grid_of(k_trace_count_wide<true>, sm_count, smem, &unused)) != cudaSuccess) return e;
  g->smem = smem;
  return cudaSuccess;
}

cudaError_t wide_launch_batch(bool any, int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* rays,
                              unsigned long long n, TutuHit* out, uint8_t* out_any, unsigned long long* next, const unsigned* perm) {
  if (any)
    k_trace_wide<true><<<grid, kTraceBlock, smem, s>>>(sc, rays, n, nullptr, out_any, next, perm);
  else
    k_trace_wide<false><<<grid, kTraceBlock, smem, s>>>(sc, rays, n, out, nullptr, next, perm);
  return cudaGetLastError();
}

cudaError_t wide_launch_count(bool any, int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* rays,
                              unsigned long long n, unsigned long long* counts) {
  if (any)
    k_trace_count_wide<true><<<grid, kTraceBlock, smem, s>>>(sc, rays, n, counts);
  else
    k_trace_count_wide<false><<<grid, kTraceBlock, smem, s>>>(sc, rays, n, counts);
  return cudaGetLastError();
}

cudaError_t wide_launch_wf_extend(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const WfBuffers& b, int cur) {
  wf_extend_wide<<<grid, kTraceBlock, smem, s>>>(sc, b, cur);
  return cudaGetLastError();
}

cudaError_t wide_launch_wf_shadow(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const WfBuffers& b, int nxt) {
  wf_shadow_wide<<<grid, kTraceBlock, smem, s>>>(sc, b, nxt);
  return cudaGetLastError();
}

cudaError_t wide_launch_q_extend(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* ro, const float4* rd,
                                 float4* hit, const unsigned* n_ptr, unsigned long long* cursor) {
  q_extend_wide<<<grid, kTraceBlock, smem, s>>>(sc, ro, rd, hit, n_ptr, cursor);
  return cudaGetLastError();
}

cudaError_t wide_launch_q_shadow_add(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* so,
                                     const float4* sd, const float4* scn, float* accum, const unsigned* n_ptr,
                                     unsigned long long* cursor) {
  q_shadow_add_wide<<<grid, kTraceBlock, smem, s>>>(sc, so, sd, scn, accum, n_ptr, cursor);
  return cudaGetLastError();
}

}  // namespace tutu
