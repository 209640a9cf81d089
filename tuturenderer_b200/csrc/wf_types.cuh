// wf_types.cuh — queue records of the wavefront path tracer shared by the shading kernels (wavefront.cuh) and
// the wide-tree queue tracers (trace_kernels.cu, a separate translation unit).
#pragma once
#include <cstddef>
#include "shade.cuh"

namespace tutu {

constexpr uint32_t kShadowFinalDst = 0xFFFFFFFFu;  // sh_d.w of a shadow ray whose path has already ended (= vertex.cuh kShadowFinal)
constexpr uint32_t kDeadQueueEntry = 0xFFFFFFFFu;  // BDPT walk queue: q_d.w of an entry to skip (= bdpt.cuh kDeadEntry)

// wf_shade reserves queue space in chunks (wavefront.cuh: "queue appends"), so a queue may hold a few entries that no
// path was written to: the unused end of a block's last chunk.  The block marks them before it exits — ray_o.w (the
// roulette number) / sh_o.w (the distance) = kDeadQueueEntry — the tracers skip them, wf_extend* writes kDeadSlot into
// their hit record so that wf_classify / wf_shade skip them too, and WfCtl::dead_* keeps the ray statistics exact.
#ifndef TUTU_APPEND_ITERS
#define TUTU_APPEND_ITERS 8
#endif
constexpr unsigned kAppendIters = TUTU_APPEND_ITERS;  // a block's reserve per queue = this x blockDim entries; 0 = one atomic per block iteration
constexpr int kDeadSlot = -2;                          // Hit::slot of a dead queue entry (-1 = miss)

struct WfCtl {
  unsigned n_cur;
  unsigned done;
  // n_next (low word) and n_shadow (high word) are bumped by ONE 64-bit atomicAdd per block of wf_shade
  unsigned n_next;
  unsigned n_shadow;
  unsigned long long next_path;
  unsigned long long total_paths;
  unsigned long long sum_extend;
  unsigned long long sum_shadow;
  unsigned long long nan_samples;
  unsigned long long iterations;
  unsigned long long cursor_extend;  // ray-queue cursors of the persistent tracers
  unsigned long long cursor_shadow;
  unsigned class_count[8];  // wf_classify: queue entries per shading class (kShadeClasses)
  // dead entries (see kAppendIters) of the current path queue, of the one being written and of the shadow queue
  unsigned dead_cur, dead_next, dead_shadow, reserved;
};
static_assert(offsetof(WfCtl, n_next) % 8 == 0 && offsetof(WfCtl, n_shadow) == offsetof(WfCtl, n_next) + 4,
              "n_next/n_shadow must form one aligned 64-bit word");

struct WfBuffers {
  // path queues, [2] = ping-pong
  float4* ray_o[2];  // o.xyz, roulette number of the vertex that spawned the ray (Philox slot 5 of its depth)
  float4* ray_d[2];  // d.xyz, q = 2 (o - x_prev) . d, the cross term of |x_hit - x_prev|^2 (shade_vertex)
  float4* st0[2];    // beta.xyz, bits(pixel)
  float4* st1[2];    // tp.xyz, bits(sample)
  float4* st2[2];    // L.xyz, bits(depth | mode<<8 | flags)
  float4* st3[2];    // f_r*cos_theta of the previous vertex .xyz, mat_pdf
  float4* hit;       // t, u, v, bits(slot code)
  // shadow queue
  float4* sh_o;  // o.xyz, dist
  float4* sh_d;  // d.xyz, bits(destination index in the next path queue | kShadowFinal)
  float4* sh_c;  // beta * NEE term .xyz, bits(pixel)
  float4* sh_L;  // L.xyz of a path that already ended (only for kShadowFinal)
  WfCtl* ctl;
  float* accum;  // width*height*3 sums
  unsigned capacity;
  // shading-class order of the current queue (wf_classify), kShadeClasses lists of `capacity` entries;
  // nullptr = shade in queue order (scenes with one shading class)
  unsigned* class_perm;
};

__device__ __forceinline__ void accum_add(float* accum, WfCtl* ctl, uint32_t pixel, f3 L) {
  // PathTracing.hpp:510-511: a sample with any NaN component is dropped (still divided by SPP)
  if (any_nan(L)) {
    atomicAdd(&ctl->nan_samples, 1ull);
    return;
  }
  float* p = accum + (size_t)pixel * 3;
  atomicAdd(p + 0, L.x);
  atomicAdd(p + 1, L.y);
  atomicAdd(p + 2, L.z);
}

// warp-aggregated append: one atomicAdd per warp, lanes take consecutive slots
__device__ __forceinline__ unsigned warp_append(unsigned* counter, bool want) {
  const unsigned mask = __ballot_sync(0xFFFFFFFFu, want);
  if (mask == 0u) return 0u;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(mask) - 1;
  unsigned base = 0u;
  if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  return base + (unsigned)__popc(mask & ((1u << lane) - 1u));
}

#ifndef TUTU_PACKET_RAYS
#define TUTU_PACKET_RAYS 128
#endif
constexpr unsigned kPacketRays = TUTU_PACKET_RAYS;  // rays per queue fetch (one same-address atomic each)

}  // namespace tutu
