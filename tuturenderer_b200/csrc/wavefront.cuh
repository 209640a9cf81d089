// wavefront.cuh — the PathTracing integrator (reference include/PathTracing.hpp:136-279 traceRay
// MIS branch, :80-134 calcForRefractive, :352-475/:485-516 integrate + sub_render_pt) as a
// wavefront pipeline:
//
//   raygen  -> extend (closest hit) -> shade (arrival MIS / roulette of the previous vertex,
//              material, NEE sample, BSDF sample) -> shadow (any hit, adds the NEE term)
//
// The reference's recursion is tail-like (value = direct_k + coe_k * value_{k+1}); it becomes an
// iteration carrying two throughputs: beta (product of every coe) and tp (the roulette tracker,
// reset to 1 while depth <= MIN_DEPTH and after refractive vertices).  A path's radiance L is
// kept per path and added to the frame buffer once, so the reference's per-sample NaN filter
// (PathTracing.hpp:510) keeps its meaning.
//
// Queues live in HBM as float4 SoA, ping-pong per iteration; survivors and shadow rays are
// appended into space that each block of wf_shade reserves in chunks ("queue appends" below).  Free
// slots are refilled by raygen every iteration, so the wavefront stays full until the last samples.
#pragma once
#include <cstddef>
#include <cuda.h>  // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "vertex.cuh"
#include "wf_types.cuh"

namespace tutu {

// ---- control ---------------------------------------------------------------------------------
__global__ void wf_ctl_after_raygen(WfCtl* ctl, unsigned capacity) {
  const unsigned free_slots = capacity > ctl->n_cur ? capacity - ctl->n_cur : 0u;  // n_cur may include dead entries
  const unsigned long long left = ctl->total_paths - ctl->next_path;
  const unsigned add = (unsigned)(left < (unsigned long long)free_slots ? left : free_slots);
  ctl->n_cur += add;
  ctl->next_path += add;
  ctl->n_next = 0;
  ctl->n_shadow = 0;
  ctl->cursor_extend = 0;
  ctl->cursor_shadow = 0;
  ctl->dead_next = 0;
  ctl->dead_shadow = 0;
  ctl->sum_extend += ctl->n_cur - ctl->dead_cur;
  for (int c = 0; c < 8; ++c) ctl->class_count[c] = 0;
}
__global__ void wf_ctl_after_iter(WfCtl* ctl) {
  if (ctl->done) return;  // the host polls every few iterations: the launches after the last one are no-ops
  ctl->sum_shadow += ctl->n_shadow - ctl->dead_shadow;
  ctl->n_cur = ctl->n_next;
  ctl->dead_cur = ctl->dead_next;
  ctl->iterations += 1;
  ctl->done = (ctl->n_cur == ctl->dead_cur && ctl->next_path >= ctl->total_paths) ? 1u : 0u;
}

// ---- raygen: PathTracing.hpp:499-509 ---------------------------------------------------------
// path g -> pixel = g % npix, sample = sample_begin + g / npix; primary rays carry no jitter, the
// pixel position is ul + x*delta_h + y*delta_v + c_off_v + c_off_v (sic), evaluated with the
// reference's operation order in exact fp32.
__global__ void __launch_bounds__(256)
wf_raygen(WfBuffers b, int cur, RayGenK k, unsigned sample_begin) {
  const WfCtl c = *b.ctl;
  const unsigned free_slots = b.capacity > c.n_cur ? b.capacity - c.n_cur : 0u;
  const unsigned long long left = c.total_paths - c.next_path;
  const unsigned add = (unsigned)(left < (unsigned long long)free_slots ? left : free_slots);
  const unsigned npix = (unsigned)k.width * (unsigned)k.height;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < add; i += gridDim.x * blockDim.x) {
    const unsigned long long g = c.next_path + i;
    const unsigned pixel = (unsigned)(g % npix);
    const unsigned sample = sample_begin + (unsigned)(g / npix);
    const float x = (float)(pixel % (unsigned)k.width), y = (float)(pixel / (unsigned)k.width);
    float px = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[0], __fmul_rn(k.dh[0], x)), __fmul_rn(k.dv[0], y)), k.cov[0]), k.cov[0]);
    float py = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[1], __fmul_rn(k.dh[1], x)), __fmul_rn(k.dv[1], y)), k.cov[1]), k.cov[1]);
    float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[2], __fmul_rn(k.dh[2], x)), __fmul_rn(k.dv[2], y)), k.cov[2]), k.cov[2]);
    float dx = __fsub_rn(px, k.eye[0]), dy = __fsub_rn(py, k.eye[1]), dz = __fsub_rn(pz, k.eye[2]);
    normalize_rn(dx, dy, dz);
    const unsigned s = c.n_cur + i;
    b.ray_o[cur][s] = make_float4(k.eye[0], k.eye[1], k.eye[2], 0.f);
    b.ray_d[cur][s] = make_float4(dx, dy, dz, 0.f);
    b.st0[cur][s] = make_float4(1.f, 1.f, 1.f, __uint_as_float(pixel));
    b.st1[cur][s] = make_float4(1.f, 1.f, 1.f, __uint_as_float(sample));
    b.st2[cur][s] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u | (kModeFresh << 8)));
  }
}

// ---- extend: IIntersectStrategy::UpdateInter -> getIntersection -------------------------------
// Ray packets pulled from the queue with one atomicAdd per warp (lane 0) and a shuffle
// broadcast: a warp that drew short rays moves on to the next packet instead of idling behind the
// slowest warp of a statically partitioned grid.
// KIND: 0 = binary trees, traversal stack in shared memory (trees too big for L1: DESIGN.md 5.10), 1 = small scene (flat
// tests), 2 = binary trees, stack in local memory; the wide-tree tracers are in trace_kernels.cu
template <int KIND>
__global__ void __launch_bounds__(256)
wf_extend(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int cur) {
  extern __shared__ unsigned long long s_stack[];  // KIND 0: traversal stack (trace.cuh: SharedStack)
  const unsigned n = b.ctl->n_cur;
  const float4* __restrict__ ro = b.ray_o[cur];
  const float4* __restrict__ rd = b.ray_d[cur];
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&b.ctl->cursor_extend, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned i = (unsigned)base + k + lane;
      if (i < n) {
        const float4 o = __ldcs(ro + i);
        const float4 d = __ldcs(rd + i);
        Hit h;
        if (__float_as_uint(o.w) == kDeadQueueEntry)
          h.t = 0.f, h.u = 0.f, h.v = 0.f, h.slot = kDeadSlot;
        else if constexpr (KIND == 1)
          traverse_small<false>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h);
        else if constexpr (KIND == 2)  // incoherent queue: per-lane walk (batched primitive tests only pay on sorted batches)
          traverse_structured<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h);
        else
          traverse_shared<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h, s_stack);
        __stcs(b.hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
      }
      __syncwarp();
    }
  }
}

// ---- shadow: isShadowRayBlocked -> hasIntersection, then the deferred NEE add ------------------
template <int KIND>
__global__ void __launch_bounds__(256)
wf_shadow(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int nxt) {
  extern __shared__ unsigned long long s_stack[];  // KIND 0: traversal stack, 32-bit entries (trace.cuh: SharedStack<true>)
  const unsigned n = b.ctl->n_shadow;
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&b.ctl->cursor_shadow, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned j = (unsigned)base + k + lane;
      float4 o = make_float4(0.f, 0.f, 0.f, __uint_as_float(kDeadQueueEntry)), d = o;
      if (j < n) o = __ldcs(b.sh_o + j), d = __ldcs(b.sh_d + j);
      if (__float_as_uint(o.w) != kDeadQueueEntry) {
        Hit h;
        bool blocked;
        if constexpr (KIND == 1)
          blocked = traverse_small<true>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h);
        else if constexpr (KIND == 2)
          blocked = traverse_structured<true>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h);
        else
          blocked = traverse_shared<true>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h, s_stack);
        const unsigned dst = __float_as_uint(d.w);
        if (dst == kShadowFinal) {
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

// ---- small scenes: two-phase flat tests (trace.cuh: SmallPark) ----------------------------------
#ifndef TUTU_SMALL_MIN_BLOCKS
#define TUTU_SMALL_MIN_BLOCKS 1
#endif
constexpr int kSmallBlock = 256;
__global__ void __launch_bounds__(kSmallBlock, TUTU_SMALL_MIN_BLOCKS)
wf_extend_small(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int cur) {
  __shared__ SmallPark s_park[kSmallBlock / 32];
  SmallPark& pk = s_park[threadIdx.x >> 5];
  const unsigned n = b.ctl->n_cur;
  const float4* __restrict__ ro = b.ray_o[cur];
  const float4* __restrict__ rd = b.ray_d[cur];
  const unsigned lane = threadIdx.x & 31u;
  unsigned parked = 0u;  // warp-uniform
  auto sink = [&](unsigned i, const Ray&, float, const Hit& h, bool) {
    __stcs(b.hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
  };
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&b.ctl->cursor_extend, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) break;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned i = (unsigned)base + k + lane;
      bool more = false;
      Ray r{};
      Hit h{};
      unsigned mask = 0u;
      if (i < n) {
        const float4 o = __ldcs(ro + i);
        const float4 d = __ldcs(rd + i);
        if (__float_as_uint(o.w) == kDeadQueueEntry) {
          h.t = 0.f, h.u = 0.f, h.v = 0.f, h.slot = kDeadSlot;
          sink(i, r, 0.f, h, false);
        } else {
          r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
          bool blocked;
          more = !small_first_pass<false>(sc, ss, r, 0.f, h, mask, blocked);
          if (!more) sink(i, r, 0.f, h, false);
        }
      }
      parked = small_park_push(pk, parked, more, r, 0.f, i, h, mask);
      while (parked >= 32u) parked = small_park_drain<false>(sc, pk, parked, false, sink);
    }
  }
  while (parked) parked = small_park_drain<false>(sc, pk, parked, parked <= 32u, sink);
}

__global__ void __launch_bounds__(kSmallBlock, TUTU_SMALL_MIN_BLOCKS)
wf_shadow_small(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int nxt) {
  __shared__ SmallPark s_park[kSmallBlock / 32];
  SmallPark& pk = s_park[threadIdx.x >> 5];
  const unsigned n = b.ctl->n_shadow;
  const unsigned lane = threadIdx.x & 31u;
  unsigned parked = 0u;
  auto sink = [&](unsigned j, const Ray&, float, const Hit&, bool blocked) {
    const unsigned dst = __float_as_uint(__ldcs(b.sh_d + j).w);
    if (dst == kShadowFinal) {
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
  };
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&b.ctl->cursor_shadow, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) break;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned j = (unsigned)base + k + lane;
      bool more = false;
      Ray r{};
      Hit h{};
      unsigned mask = 0u;
      float dis = 0.f;
      float4 o = make_float4(0.f, 0.f, 0.f, __uint_as_float(kDeadQueueEntry)), d = o;
      if (j < n) o = __ldcs(b.sh_o + j), d = __ldcs(b.sh_d + j);
      if (__float_as_uint(o.w) != kDeadQueueEntry) {
        r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
        dis = o.w;
        bool blocked;
        more = !small_first_pass<true>(sc, ss, r, dis, h, mask, blocked);
        if (!more) sink(j, r, dis, h, blocked);
      }
      parked = small_park_push(pk, parked, more, r, dis, j, h, mask);
      while (parked >= 32u) parked = small_park_drain<true>(sc, pk, parked, false, sink);
    }
  }
  while (parked) parked = small_park_drain<true>(sc, pk, parked, parked <= 32u, sink);
}

// ---- shading classes ----------------------------------------------------------------------------
// ncu on the glass / texture scene (profiles/r01_glass_shade_lanes.txt): wf_shade runs with 8.9 of 32
// lanes per instruction — the GGX / glass / texture code (1500+ instructions) executes for the one or
// two lanes of a warp that hit such a material while the Lambertian lanes wait.  On scenes with more
// than one class the queue is therefore ordered by the hit's shading class between extend and shade:
// wf_classify appends every queue index to its class's list (block-aggregated atomics, one per class and
// block iteration; warp-contiguous runs keep the Lambertian majority's loads coalesced) and wf_shade
// walks the lists back to back.  Path results do not depend on the order (every path writes its own
// continuation / frame-buffer sample), only the float summation order of the atomics changes.
constexpr int kShadeClasses = 8;  // 0 miss | 1 + MaterialType (Lambertian .. UNLIT) | 7 textured Lambertian
__device__ __forceinline__ int shade_class(const DevScene& sc, const float4 hit) {
  const int code = __float_as_int(hit.w);
  if (code == kDeadSlot) return -1;
  if (code < 0) return 0;
  const uint32_t flags = __float_as_uint(__ldg(sc.shade + 4 * (size_t)((uint32_t)code & kSlotMask) + 3).w);
  const int type = __float_as_int(__ldg(sc.materials + 4 * (size_t)(flags & 0x3FFFFFFFu)).w);
  if (type == TUTU_MAT_LAMBERTIAN && (flags & 0x80000000u)) return 7;
  return 1 + (type < 0 ? 0 : (type > 5 ? 5 : type));
}

__global__ void __launch_bounds__(256)
wf_classify(const __grid_constant__ DevScene sc, WfBuffers b) {
  const unsigned n = b.ctl->n_cur;
  __shared__ unsigned s_cnt[kShadeClasses][8];  // [class][warp] -> exclusive prefix within the block
  __shared__ unsigned s_base[kShadeClasses];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const int c = i < n ? shade_class(sc, __ldg(b.hit + i)) : -1;
    unsigned my_mask = 0u;
#pragma unroll
    for (int k = 0; k < kShadeClasses; ++k) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, c == k);
      if (lane == 0) s_cnt[k][warp] = (unsigned)__popc(m);
      if (c == k) my_mask = m;
    }
    __syncthreads();
    if (threadIdx.x < kShadeClasses) {
      unsigned total = 0;
      for (int w = 0; w < 8; ++w) {
        const unsigned v = s_cnt[threadIdx.x][w];
        s_cnt[threadIdx.x][w] = total;
        total += v;
      }
      s_base[threadIdx.x] = total ? atomicAdd(&b.ctl->class_count[threadIdx.x], total) : 0u;
    }
    __syncthreads();
    if (c >= 0)
      b.class_perm[(size_t)c * b.capacity + s_base[c] + s_cnt[c][warp] + (unsigned)__popc(my_mask & ((1u << lane) - 1u))] = i;
    __syncthreads();
  }
}

// Compiled for up to 256 threads / 2 blocks per SM (128 registers); the launch picks the block size: 64 threads for scenes
// shaded in queue order (Cornell 1024^2, Mpaths/s with 32 / 64 / 96 / 128 / 256 threads: 2049 / 2184 / 2130 / 2172 / 2073 —
// small blocks keep the one barrier of an iteration between two warps), 256 for mixed-material scenes (many small
// blocks on different material code paths thrash the instruction cache: 426 / 353 / 317 Mpaths/s with 256 / 128 / 64).
#ifndef TUTU_SHADE_MIN_BLOCKS
#define TUTU_SHADE_MIN_BLOCKS 2
#endif
#ifndef TUTU_SHADE_BLOCK
#define TUTU_SHADE_BLOCK 256
#endif
#ifndef TUTU_SHADE_BLOCK_SIMPLE
#define TUTU_SHADE_BLOCK_SIMPLE 64
#endif
constexpr int kShadeBlockSimple = TUTU_SHADE_BLOCK_SIMPLE;

// ---- TMA staging of the queue records ------------------------------------------------------------------------
// A block's records of one iteration are seven 16 B x blockDim slices of the queue arrays.  One thread asks the copy
// engine for the NEXT iteration's slices (one cp.async.bulk.tensor tile, completion counted on an mbarrier) before
// the block shades the current ones, so the HBM latency that every block iteration used to start with (ncu before the
// change: 19 % of wf_shade's stall samples on the first use of the queue loads) is overlapped with shading, and no
// registers are spent on bytes in flight.  Two stages; a third costs L1 and measured slower (2184 -> 2164 Mpaths/s).
// Scenes shaded through class lists gather their records and keep the direct loads.
#ifndef TUTU_SHADE_STAGES
#define TUTU_SHADE_STAGES 2
#endif
constexpr unsigned kShadeStages = TUTU_SHADE_STAGES;  // block iterations in flight: the current one + (stages - 1) being fetched
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok)
               : "r"(smem_addr(bar)), "r"(parity)
               : "memory");
  return ok != 0u;
}
// One tile of the queue pool: blockDim entries of 7 arrays, a box of the 3-D view {floats of a run of <= 64 entries, runs,
// arrays} that the host encodes over the lane's pool (tutu_b200.cu: encode_pool_map).  The pool is laid out
// as  set 0 (ray_o ray_d st0 st1 st2 st3) | hit | set 1 | shadow arrays,  so the seven arrays wf_shade reads are adjacent
// for either ping-pong index: rows 0..6 for cur = 0 (hit last), rows 6..12 for cur = 1 (hit first).  Entries past the end
// of the pool arrive as zeros, entries past the end of the queue are never looked at.  One instruction per block iteration;
// the seven 1-D bulk copies it replaces cost 140 warp instructions (ncu: 4.5 % of the kernel, all in warp 0).
// `run_index` = base / (entries per run) = (block iterations before this one in the whole grid) x (runs per block).
__device__ __forceinline__ void shade_stage_fetch(const CUtensorMap* map, int cur, unsigned run_index, float4* stage, unsigned long long* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy reads of this buffer are done (barrier)
  mbar_expect_tx(bar, 7u * 16u * blockDim.x);
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_addr(stage)),
               "l"(map), "r"(0), "r"((int)run_index), "r"(cur ? 6 : 0), "r"(smem_addr(bar))
               : "memory");
}

// ---- queue appends -----------------------------------------------------------------------------------------------
// Survivors and shadow rays are appended to two queues whose counters live in one 64-bit word.  Round 1 took one
// atomicAdd per block iteration (after per-warp atomics had put half of the kernel's stall samples on the shuffle that
// waits for them); the block then sat at a barrier for the L2 round trip of that atomic once per iteration (ncu,
// profiles/r02_steady_shade_stalls.txt: 21 % of the stall samples on the barriers of a 256-thread block).  Now a block
// keeps a reserve of kAppendIters x blockDim entries per queue and takes its entries from the reserve: the global atomic — and
// the second barrier that publishes its result — happen only in the iterations whose appends do not fit (about one in
// eight to twelve).  The appends of such an iteration fill the old reserve to its last entry and continue in the new
// one, so a queue has no holes except the unused end of each block's LAST reserve; a block's requests shrink as it
// runs out of iterations (its last ones ask for exactly what they need), it writes dead markers (wf_types.cuh) over
// what is left when it exits and adds the count to WfCtl::dead_*.
struct AppendState {
  unsigned lo[2], rem[2];  // [path queue, shadow queue]: next free entry of the reserve, entries left in it
};

template <int SPEC>
__device__ __forceinline__ void wf_shade_body(const DevScene& sc, const WfBuffers& b, const CUtensorMap* pool_map, int cur,
                                              uint64_t seed, unsigned (*s_cnt)[TUTU_SHADE_BLOCK / 32], AppendState* s_app,
                                              unsigned* s_new, unsigned* s_pref) {
  const int nxt = cur ^ 1;
  const unsigned n = b.ctl->n_cur;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned n_warps = blockDim.x >> 5;
  // class lists (wf_classify): entry j of the concatenated lists -> queue index
  unsigned n_listed = n;
  if (b.class_perm) {
    if (threadIdx.x == 0) {
      unsigned acc = 0u;
      for (int c = 0; c < kShadeClasses; ++c) {
        s_pref[c] = acc;
        acc += b.ctl->class_count[c];
      }
      s_pref[kShadeClasses] = acc;
    }
    __syncthreads();
    n_listed = s_pref[kShadeClasses];  // dead queue entries are in no list
  }
  extern __shared__ __align__(128) float4 s_stage[];  // [stages][7 arrays][blockDim]; none for class-list scenes
  __shared__ unsigned long long s_full[kShadeStages];
  const bool piped = b.class_perm == nullptr;          // class lists gather their records: direct loads
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned runs_per_block = blockDim.x / min(blockDim.x, 64u);  // encode_pool_map: a run is min(block, 64) entries
  if (threadIdx.x == 0) s_app[0] = AppendState{{0u, 0u}, {0u, 0u}};
  if (piped) {
    if (threadIdx.x == 0) {
      for (unsigned k = 0; k < kShadeStages; ++k) mbar_init(&s_full[k], 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (unsigned k = 0; k + 1 < kShadeStages; ++k) {  // prologue: the first stages - 1 iterations
        const unsigned at = blockIdx.x * blockDim.x + k * stride;
        if (at < n) shade_stage_fetch(pool_map, cur, (blockIdx.x + k * gridDim.x) * runs_per_block, s_stage + k * 7u * blockDim.x, &s_full[k]);
      }
  }
  const unsigned first = blockIdx.x * blockDim.x;
  const unsigned n_iters = first < n_listed ? (n_listed - 1u - first) / stride + 1u : 0u;  // block iterations of this block
  unsigned iter = 0u;
  for (unsigned base = first; iter < n_iters; base += stride, ++iter) {
    const unsigned j = base + threadIdx.x;
    const bool valid = j < n_listed;
    unsigned i = j;
    const unsigned slot = iter % kShadeStages;
    const float4* stage = s_stage + slot * 7u * blockDim.x;
    if (piped) {
      // fetch iteration iter + stages - 1 into the buffer that iteration iter - 1 read (every thread has passed that
      // iteration's barrier), then wait for this iteration's records
      const unsigned ahead = (iter + kShadeStages - 1u) % kShadeStages;
      const unsigned long long next = (unsigned long long)base + (unsigned long long)(kShadeStages - 1u) * stride;
      if (threadIdx.x == 0 && next < n)
        shade_stage_fetch(pool_map, cur, (blockIdx.x + (iter + kShadeStages - 1u) * gridDim.x) * runs_per_block, s_stage + ahead * 7u * blockDim.x,
                          &s_full[ahead]);
      while (!mbar_try_wait(&s_full[slot], (iter / kShadeStages) & 1u)) {
      }
    }
    if (b.class_perm && valid) {
      unsigned c = 0u, start = 0u;
#pragma unroll
      for (int k = 1; k < kShadeClasses; ++k) {
        const unsigned p = s_pref[k];
        if (j >= p) c = (unsigned)k, start = p;
      }
      i = b.class_perm[(size_t)c * b.capacity + (j - start)];
    }
    ShadeOut out;
    out.cont = out.shadow = out.finished = false;
    f3 L = mk(0.f);
    uint32_t pixel = 0;
    float4 s1 = make_float4(0, 0, 0, 0);
    if (valid) {
      // queue records: from the staging buffer, or (class lists) straight from HBM — touched once per iteration, so
      // streamed past L1/L2 residency (.cs) to keep the scene tables cached.  st3 only matters to a vertex reached
      // through x_inter, but it is fetched with the others: asking for it after s2 has told the mode would put a
      // second memory round trip on the critical path.
      float4 o, d, s0, s2, hit, s3;
      if (piped) {
        const unsigned bd = blockDim.x;
        const float4* rec = stage + (cur ? bd : 0u) + threadIdx.x;  // tile rows: cur = 0: set 0 then hit; cur = 1: hit then set 1
        o = rec[0], d = rec[bd], s0 = rec[2 * bd], s1 = rec[3 * bd], s2 = rec[4 * bd], s3 = rec[5 * bd];
        hit = stage[(cur ? 0u : 6u * bd) + threadIdx.x];
      } else {
        o = __ldcs(b.ray_o[cur] + i), d = __ldcs(b.ray_d[cur] + i), s0 = __ldcs(b.st0[cur] + i), s1 = __ldcs(b.st1[cur] + i);
        s2 = __ldcs(b.st2[cur] + i), hit = __ldcs(b.hit + i), s3 = __ldcs(b.st3[cur] + i);
      }
      if (__float_as_int(hit.w) != kDeadSlot) {
        const uint32_t dm = __float_as_uint(s2.w);
        const uint32_t depth = dm & 0xFFu, mode = (dm >> 8) & 1u;
        pixel = __float_as_uint(s0.w);
        L = mk(s2.x, s2.y, s2.z);
        Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
        shade_vertex<SPEC>(sc, seed, r, hit, pixel, __float_as_uint(s1.w), depth, mode, dm, mk(s0.x, s0.y, s0.z),
                     mk(s1.x, s1.y, s1.z), L, s3, d.w, o.w, out);
      }
    }
    // queue appends (above): warp ballots -> counts in shared memory -> every warp sums the counts of the warps before it
    const unsigned par = iter & 1u;
    const unsigned cmask = __ballot_sync(0xFFFFFFFFu, out.cont);
    const unsigned smask = __ballot_sync(0xFFFFFFFFu, out.shadow);
    const unsigned mine = (unsigned)__popc(cmask) | ((unsigned)__popc(smask) << 16);  // both counts in one word (<= 1024 each)
    if (lane == 0) s_cnt[par][warp] = mine;
    __syncthreads();  // the one barrier of an iteration whose appends fit the reserve; it also ends the reads of `stage`
    constexpr unsigned kMaxWarps = TUTU_SHADE_BLOCK / 32;
    unsigned scan = lane < n_warps ? s_cnt[par][lane] : 0u;  // inclusive prefix over the warps, by shuffles
#pragma unroll
    for (unsigned o = 1; o < kMaxWarps; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xFFFFFFFFu, scan, o);
      if (lane >= o) scan += t;
    }
    const unsigned total = __shfl_sync(0xFFFFFFFFu, scan, kMaxWarps - 1), before = __shfl_sync(0xFFFFFFFFu, scan, warp) - mine;
    const unsigned tc = total & 0xFFFFu, ts = total >> 16;
    unsigned pc = before & 0xFFFFu, ps = before >> 16;
    const AppendState st = s_app[par];
    const bool need_c = tc > st.rem[0], need_s = ts > st.rem[1];
    unsigned new_c = 0u, new_s = 0u;
    // the reserve a block asks for shrinks with the iterations it has left (half of what they could append at most), so
    // what is left over when it exits stays small whatever the size of the queue
    const unsigned left = n_iters - 1u - iter;
    const unsigned keep = min(kAppendIters, left / 2u) * blockDim.x;
    if (need_c | need_s) {  // block-uniform
      if (threadIdx.x == 0) {  // both counters in one atomic: {n_shadow : n_next}
        const unsigned long long want = ((unsigned long long)(need_s ? ts - st.rem[1] + keep : 0u) << 32) |
                                        (unsigned long long)(need_c ? tc - st.rem[0] + keep : 0u);
        const unsigned long long got = atomicAdd(reinterpret_cast<unsigned long long*>(&b.ctl->n_next), want);
        s_new[0] = (unsigned)got;
        s_new[1] = (unsigned)(got >> 32);
      }
      __syncthreads();
      new_c = s_new[0], new_s = s_new[1];
    }
    if (threadIdx.x == 0) {
      AppendState nx;
      nx.lo[0] = need_c ? new_c + (tc - st.rem[0]) : st.lo[0] + tc;
      nx.rem[0] = need_c ? keep : st.rem[0] - tc;
      nx.lo[1] = need_s ? new_s + (ts - st.rem[1]) : st.lo[1] + ts;
      nx.rem[1] = need_s ? keep : st.rem[1] - ts;
      s_app[par ^ 1u] = nx;  // read after the next iteration's barrier
    }
    const unsigned lt = (1u << lane) - 1u;
    pc += (unsigned)__popc(cmask & lt), ps += (unsigned)__popc(smask & lt);
    const unsigned ci = pc < st.rem[0] ? st.lo[0] + pc : new_c + (pc - st.rem[0]);
    const unsigned si = ps < st.rem[1] ? st.lo[1] + ps : new_s + (ps - st.rem[1]);
    if (out.cont) {
      __stcs(b.ray_o[nxt] + ci, make_float4(out.o.x, out.o.y, out.o.z, out.rr_u));
      __stcs(b.ray_d[nxt] + ci, make_float4(out.d.x, out.d.y, out.d.z, out.q));
      __stcs(b.st0[nxt] + ci, make_float4(out.beta.x, out.beta.y, out.beta.z, __uint_as_float(pixel)));
      __stcs(b.st1[nxt] + ci, make_float4(out.tp.x, out.tp.y, out.tp.z, s1.w));
      // st2 is read-modify-written by wf_shadow right after: keep it in L2 (default policy)
      b.st2[nxt][ci] = make_float4(L.x, L.y, L.z, __uint_as_float(out.depth_mode));
      if (((out.depth_mode >> 8) & 1u) == kModeXInter) {
        __stcs(b.st3[nxt] + ci, make_float4(out.fcos.x, out.fcos.y, out.fcos.z, out.mat_pdf));
      }
    }
    if (out.shadow) {
      b.sh_o[si] = make_float4(out.so.x, out.so.y, out.so.z, out.sdist);
      b.sh_d[si] = make_float4(out.sd.x, out.sd.y, out.sd.z, __uint_as_float(out.cont ? ci : kShadowFinal));
      b.sh_c[si] = make_float4(out.sc.x, out.sc.y, out.sc.z, __uint_as_float(pixel));
      if (!out.cont) b.sh_L[si] = make_float4(L.x, L.y, L.z, 0.f);
    } else if (out.finished) {
      accum_add(b.accum, b.ctl, pixel, L);
    }
  }
  // what is left of the reserves: dead entries
  __syncthreads();
  const AppendState st = s_app[iter & 1u];
  for (unsigned t = threadIdx.x; t < st.rem[0]; t += blockDim.x)
    b.ray_o[nxt][st.lo[0] + t] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kDeadQueueEntry));
  for (unsigned t = threadIdx.x; t < st.rem[1]; t += blockDim.x)
    b.sh_o[st.lo[1] + t] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kDeadQueueEntry));
  if (threadIdx.x == 0) {
    if (st.rem[0]) atomicAdd(&b.ctl->dead_next, st.rem[0]);
    if (st.rem[1]) atomicAdd(&b.ctl->dead_shadow, st.rem[1]);
  }
}

__global__ void __launch_bounds__(TUTU_SHADE_BLOCK, TUTU_SHADE_MIN_BLOCKS)
wf_shade(const __grid_constant__ DevScene sc, const __grid_constant__ CUtensorMap pool_map, WfBuffers b, int cur, uint64_t seed) {
  __shared__ unsigned s_cnt[2][TUTU_SHADE_BLOCK / 32];
  __shared__ AppendState s_app[2];
  __shared__ unsigned s_new[2];
  __shared__ unsigned s_pref[kShadeClasses + 1];
  wf_shade_body<0>(sc, b, &pool_map, cur, seed, s_cnt, s_app, s_new, s_pref);
}
// ---- finalize: color = estimate * SPP_inv (PathTracing.hpp:513) -------------------------------
__global__ void wf_finalize(const float* __restrict__ accum, float inv_spp, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = accum[i] * inv_spp;
}


// ---- output stage: PPMGenerator::writePixel (PPMGenerator.hpp:812-845) ---------------------------
// (int)(255 * pow(clamp(0,1,c), gamma)).  The reference's powf (glibc) is correctly rounded in all
// but a vanishing fraction of inputs; the double-precision pow rounded to float reproduces it.
__device__ __forceinline__ unsigned char quantize_channel(float c, float gamma) {
  // std::max(lo, std::min(hi, v)): NaN -> hi.  Spelled out: nvcc turns the two selects into a
  // saturate, which sends NaN to 0.
  const float cl = isnan(c) ? 1.f : fminf(fmaxf(c, 0.f), 1.f);
  float v;
  if (gamma > 0.f)
    v = __fmul_rn(255.f, (float)pow((double)cl, (double)gamma));
  else
    v = __fmul_rn(255.f, cl);
  return (unsigned char)(int)v;
}
__global__ void __launch_bounds__(256)
k_quantize(const float* __restrict__ rgb, size_t n_values, float gamma, unsigned char* __restrict__ out) {
  // 4 values (one 16-byte load, one 4-byte store) per thread and iteration
  const size_t n4 = n_values / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(rgb) + i);
    uchar4 q;
    q.x = quantize_channel(v.x, gamma), q.y = quantize_channel(v.y, gamma);
    q.z = quantize_channel(v.z, gamma), q.w = quantize_channel(v.w, gamma);
    reinterpret_cast<uchar4*>(out)[i] = q;
  }
  if (blockIdx.x == 0 && threadIdx.x < n_values % 4) {
    const size_t i = n4 * 4 + threadIdx.x;
    out[i] = quantize_channel(rgb[i], gamma);
  }
}

}  // namespace tutu
