// device_bvh.hpp — GPU build of the traversal tree for regular rays (device_bvh.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tutu {

// Linear BVH over n >= 2 leaves.  d_leaf_box: n x {lo.xyz, hi.xyz} (exact leaf boxes, by DFS slot of the reference
// tree); d_leaf_code: n x (slot | sphere bit); root box = union of all leaf boxes.  Writes n - 1 InnerNode records
// (64 B, tutu_internal.hpp) to d_inner_out, root = node 0, inner boxes = exact fmin/fmax unions, and the number
// of inner nodes on the longest root-to-leaf path to *depth_out.  Synchronises `s` before returning.
// d_scratch: device_build_lbvh_scratch_bytes(n) bytes of device memory (256-byte aligned), owned by the caller.
size_t device_build_lbvh_scratch_bytes(uint32_t n);
cudaError_t device_build_lbvh(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, const float root_lo[3],
                              const float root_hi[3], void* d_inner_out, uint32_t* depth_out, int sm_count, void* d_scratch,
                              cudaStream_t s);

// The same interface for a tree built by parallel locally-ordered clustering (PLOC): Morton-sorted leaves are merged
// bottom-up, nearest neighbour by surface area within a window, until one cluster is left.  iterations_out: merge passes.
size_t device_build_ploc_scratch_bytes(uint32_t n);
cudaError_t device_build_ploc(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, const float root_lo[3],
                              const float root_hi[3], void* d_inner_out, uint32_t* depth_out, uint32_t* iterations_out, int sm_count,
                              void* d_scratch, cudaStream_t s);

// Top-down binned SAH with the host builder's split rule (host_scene.cpp: FastBuilder): the same tree, node for node,
// wherever the host does not fall back to halving a node by its current record order.  max_depth: the host's depth
// budget (kFastTreeMaxDepth).  levels_out: level-synchronous passes over the nodes of more than 64 primitives.
size_t device_build_sah_scratch_bytes(uint32_t n);
cudaError_t device_build_sah(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, int max_depth, void* d_inner_out,
                             uint32_t* depth_out, uint32_t* levels_out, int sm_count, void* d_scratch, cudaStream_t s);

}  // namespace tutu
