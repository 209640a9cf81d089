// trace_kernels.hpp — host-callable launchers of the wide-tree walking kernels (trace_kernels.cu).  They live
// in their own translation unit: cicc 12.9 crashes on tutu_b200.cu when the wide walk is instantiated next to
// the shading kernels, and the split keeps "kernels that walk trees" apart from "kernels that shade" anyway.
#pragma once
#include <cuda_runtime.h>

#include "tutu_internal.hpp"
#include "wf_types.cuh"

namespace tutu {

constexpr int kTraceBlock = 256;

struct WideGrids {  // resident blocks per SM x SMs of each kernel for a given stack size
  int batch_closest = 0, batch_any = 0, wf_extend = 0, wf_shadow = 0, q_extend = 0, q_shadow = 0;
  size_t smem = ~(size_t)0;  // the stack size these were computed for
};
// Sets the dynamic shared-memory attribute of every kernel and fills the grids.
cudaError_t wide_grids(int sm_count, size_t stack_smem, WideGrids* out);

// ray batches (tutu_trace_closest / tutu_trace_any): rays taken in `perm` order when given
cudaError_t wide_launch_batch(bool any, int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* rays,
                              unsigned long long n, TutuHit* out, uint8_t* out_any, unsigned long long* next, const unsigned* perm);
// visit counters: counts[0] += wide nodes fetched, counts[1] += primitive tests
cudaError_t wide_launch_count(bool any, int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* rays,
                              unsigned long long n, unsigned long long* counts);
// wavefront queues: extend writes b.hit for queue `cur`; shadow adds the deferred NEE terms into queue `nxt`
cudaError_t wide_launch_wf_extend(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const WfBuffers& b, int cur);
cudaError_t wide_launch_wf_shadow(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const WfBuffers& b, int nxt);
// BDPT queues
cudaError_t wide_launch_q_extend(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* ro, const float4* rd,
                                 float4* hit, const unsigned* n_ptr, unsigned long long* cursor);
cudaError_t wide_launch_q_shadow_add(int grid, size_t smem, cudaStream_t s, const DevScene& sc, const float4* so,
                                     const float4* sd, const float4* scn, float* accum, const unsigned* n_ptr,
                                     unsigned long long* cursor);

}  // namespace tutu
