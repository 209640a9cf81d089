// resident.cuh — the PathTracing integrator for scenes whose whole geometry sits in the kernel's
// constant bank (SmallScene, <= 32 primitives: the Cornell box of BASELINE.json configs[2]) as ONE
// persistent kernel with the path state in registers.
//
// Why a second pipeline: on such a scene the wavefront's three kernels are bound by different things
// (profiles/r01c_wavefront_full.txt): wf_extend / wf_shadow issue 84 % of their slots on the flat slab +
// triangle tests, wf_shade moves ~330 B of queue records per vertex at 24 % occupancy and issues 41 %.
// Summed over a step the SMs issue ~48 % of their slots.  There is no traversal stack and no node
// fetch to hide here, so nothing needs the queues: a thread can carry its path from the camera to its
// end (PathTracing.hpp:136-279 is a tail recursion) and start the next sample the moment it ends.
//
//   work item   = one pixel x `chunk` consecutive samples, drawn by the lane from a global cursor;
//   loop body   = [regenerate if the path ended] -> closest hit (traverse_small) -> shade_vertex (the
//                 same function the wavefront calls) -> any hit for the NEE ray -> keep or end;
//   frame buffer= per item the lane sums its samples in registers in sample order (the reference's
//                 `estimate = estimate + res`, PathTracing.hpp:507-511, NaN filter per sample) and
//                 adds the sum once (3 atomics per `chunk` paths).
//
// Every path draws the Philox slots (seed; pixel, sample, depth) it draws in the wavefront and runs the
// same shade_vertex / flat tests, so the two pipelines agree to float noise (frame-buffer summation order;
// FMA contraction is decided per translation unit).  HBM traffic: the 12 MB frame buffer.
//
// Measured on a B200 (Cornell 1024^2, tools/gpu_resident.py): 1313 Mpaths/s against the wavefront's 1950.
// ncu (profiles/r01e_resident_*): 128 registers -> 16 warps per SM, 29 % of the stall samples are
// instruction-cache misses (flat tests + shading = one 10 000-instruction loop that warps run out of
// phase), 17 of 32 lanes per instruction (a lane cannot be parked and refilled the way a queue entry
// can).  What it wins is latency (one launch, no queue pools): 3.4x faster at 64x64 @ 16 spp, even at ~1 M
// paths — the automatic pipeline choice takes it below 768 Ki paths (tutu_b200.cu: wf_render).
#pragma once
#include "vertex.cuh"

namespace tutu {

#ifndef TUTU_RESIDENT_BLOCK
#define TUTU_RESIDENT_BLOCK 128
#endif
#ifndef TUTU_RESIDENT_MIN_BLOCKS
#define TUTU_RESIDENT_MIN_BLOCKS 4
#endif

struct ResidentCtl {
  unsigned long long cursor;  // next work item
  unsigned long long sum_extend;
  unsigned long long sum_shadow;
  unsigned long long nan_samples;
  unsigned long long iterations;  // loop trips of the longest-running warp (diagnostic)
};

struct ResidentArgs {
  RayGenK rk;
  unsigned sample_begin, sample_count;
  unsigned chunk;  // samples per work item
  unsigned long long n_items;
  unsigned long long seed;
  float* accum;
  ResidentCtl* ctl;
};

// host side (resident.cu): resident blocks per SM x SMs, and the launch
cudaError_t pt_resident_grid(int sm_count, int* grid);
cudaError_t pt_resident_launch(int grid, cudaStream_t s, const DevScene& sc, const SmallScene& ss,
                               const ResidentArgs& a);

#ifdef TUTU_RESIDENT_IMPL
// primary ray of a pixel, PathTracing.hpp:499-504 (same arithmetic as wf_raygen)
__device__ __forceinline__ void primary_dir(const RayGenK& k, unsigned pixel, float& dx, float& dy, float& dz) {
  const float x = (float)(pixel % (unsigned)k.width), y = (float)(pixel / (unsigned)k.width);
  const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[0], __fmul_rn(k.dh[0], x)), __fmul_rn(k.dv[0], y)), k.cov[0]), k.cov[0]);
  const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[1], __fmul_rn(k.dh[1], x)), __fmul_rn(k.dv[1], y)), k.cov[1]), k.cov[1]);
  const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(k.ul[2], __fmul_rn(k.dh[2], x)), __fmul_rn(k.dv[2], y)), k.cov[2]), k.cov[2]);
  dx = __fsub_rn(px, k.eye[0]), dy = __fsub_rn(py, k.eye[1]), dz = __fsub_rn(pz, k.eye[2]);
  normalize_rn(dx, dy, dz);
}

__global__ void __launch_bounds__(TUTU_RESIDENT_BLOCK, TUTU_RESIDENT_MIN_BLOCKS)
pt_resident(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss,
            const __grid_constant__ ResidentArgs a) {
  const unsigned npix = (unsigned)a.rk.width * (unsigned)a.rk.height;
  const unsigned s_last = a.sample_begin + a.sample_count;
  // work item
  unsigned pixel = 0, s_next = 0, s_end = 0;
  bool have_item = false, exhausted = false;
  f3 Lsum = mk(0.f);
  // path
  bool alive = false;
  f3 o = mk(0.f), d = mk(0.f), beta = mk(1.f), tp = mk(1.f), L = mk(0.f), fcos = mk(0.f);
  float mat_pdf = 0.f, rr_u = 0.f, q = 0.f;
  uint32_t dm = 0u, sample = 0u;
  unsigned n_ext = 0, n_sh = 0, n_nan = 0, trips = 0;

  for (;;) {
    if (!alive && !exhausted) {
      if (s_next == s_end) {
        if (have_item) {
          float* p = a.accum + (size_t)pixel * 3;
          atomicAdd(p + 0, Lsum.x);
          atomicAdd(p + 1, Lsum.y);
          atomicAdd(p + 2, Lsum.z);
          have_item = false;
        }
        const unsigned long long item = atomicAdd(&a.ctl->cursor, 1ull);
        if (item >= a.n_items) {
          exhausted = true;
        } else {
          pixel = (unsigned)(item % npix);
          s_next = a.sample_begin + (unsigned)(item / npix) * a.chunk;
          s_end = s_next + a.chunk < s_last ? s_next + a.chunk : s_last;
          Lsum = mk(0.f);
          have_item = true;
        }
      }
      if (!exhausted) {
        o = mk(a.rk.eye[0], a.rk.eye[1], a.rk.eye[2]);
        primary_dir(a.rk, pixel, d.x, d.y, d.z);
        beta = mk(1.f), tp = mk(1.f), L = mk(0.f), fcos = mk(0.f);
        mat_pdf = rr_u = q = 0.f;
        dm = 0u | (kModeFresh << 8);
        sample = s_next++;
        alive = true;
      }
    }
    if (__all_sync(0xFFFFFFFFu, !alive)) break;  // only when every lane has drained the cursor
    ++trips;
    if (alive) {
      const Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
      Hit h;
      traverse_small<false>(sc, ss, r, 0.f, h);
      ++n_ext;
      ShadeOut out;
      shade_vertex<0>(sc, a.seed, r, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)), pixel, sample, dm & 0xFFu,
                      (dm >> 8) & 1u, dm, beta, tp, L, make_float4(fcos.x, fcos.y, fcos.z, mat_pdf), q, rr_u, out);
      if (out.shadow) {
        Hit hs;
        ++n_sh;
        if (!traverse_small<true>(sc, ss, Ray{out.so.x, out.so.y, out.so.z, out.sd.x, out.sd.y, out.sd.z}, out.sdist, hs))
          L = L + out.sc;
      }
      if (out.cont) {
        o = out.o, d = out.d, beta = out.beta, tp = out.tp, fcos = out.fcos;
        mat_pdf = out.mat_pdf, rr_u = out.rr_u, q = out.q, dm = out.depth_mode;
      } else {
        // PathTracing.hpp:510-511: a sample with any NaN component is dropped (still divided by SPP)
        if (any_nan(L))
          ++n_nan;
        else
          Lsum = Lsum + L;
        alive = false;
      }
    }
  }
  // counters: one atomic per warp and counter
  for (int s = 16; s > 0; s >>= 1) {
    n_ext += __shfl_xor_sync(0xFFFFFFFFu, n_ext, s);
    n_sh += __shfl_xor_sync(0xFFFFFFFFu, n_sh, s);
    n_nan += __shfl_xor_sync(0xFFFFFFFFu, n_nan, s);
  }
  if ((threadIdx.x & 31u) == 0u) {
    atomicAdd(&a.ctl->sum_extend, (unsigned long long)n_ext);
    atomicAdd(&a.ctl->sum_shadow, (unsigned long long)n_sh);
    if (n_nan) atomicAdd(&a.ctl->nan_samples, (unsigned long long)n_nan);
    atomicMax(&a.ctl->iterations, (unsigned long long)trips);
  }
}
#endif  // TUTU_RESIDENT_IMPL

}  // namespace tutu
