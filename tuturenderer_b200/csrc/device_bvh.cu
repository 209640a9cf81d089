// device_bvh.cu — the traversal tree for regular rays built ON THE GPU.
//
// What may be built freely: for a ray whose three 1/d are finite the reference's answer depends only on the leaf
// boxes and the primitives (host_scene.cpp, "traversal tree for regular rays"), so any binary tree over the same
// leaves whose inner boxes are exact fmin/fmax unions returns bit-identical hits.  The host builds a binned-SAH
// tree (0.3-0.7 s for 10^6 primitives); this file builds a linear BVH in a few milliseconds:
//
//   1. lbvh_keys       48-bit Morton code of every leaf-box centroid inside the scene box (16 bits per axis)
//   2. radix_*         stable LSD radix sort of (key, leaf) pairs, 4 bits per pass, 12 passes; every thread owns 16
//                      consecutive pairs, so stability needs no intra-warp ranking: per-thread digit counts ->
//                      block scan per digit -> one global scan of the [digit][block] table -> ordered scatter
//   3. lbvh_hierarchy  Karras, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees" (HPG
//                      2012): inner node i covers a range of the sorted leaves found from the common-prefix
//                      lengths delta(i, j) (equal keys fall back to the index, so duplicates split evenly)
//   4. lbvh_refit      leaves walk to the root; the second arrival at a node owns both child boxes and goes on.
//                      Boxes are fmin/fmax of floats: exact and independent of the arrival order
//   5. lbvh_depth      deepest leaf (the traversal stacks are sized by it)
//
// Output = the 64-byte two-child nodes the binary walk reads (tutu_internal.hpp: InnerNode), root = node 0, leaf
// refs = ~(DFS slot | sphere bit) of the REFERENCE tree's leaf numbering, which is what the equal-t tie rule compares.
// No library sort: every kernel here is this repository's.  sm_100a only.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "device_bvh.hpp"

namespace tutu {
namespace {

constexpr int kSortBlock = 256;
constexpr int kSortItems = 16;                          // pairs per thread, consecutive
constexpr int kSortTile = kSortBlock * kSortItems;      // pairs per block
constexpr int kRadixBits = 4, kRadix = 1 << kRadixBits;
constexpr int kKeyBits = 48;

__device__ __forceinline__ unsigned long long spread16(unsigned v) {  // 16 bits -> every third bit of 48
  unsigned long long x = v & 0xFFFFull;
  x = (x | (x << 32)) & 0x00FF00000000FFFFull;  // not needed for 16 bits, kept for the general pattern
  x = (x | (x << 16)) & 0x00FF0000FF0000FFull;
  x = (x | (x << 8)) & 0xF00F00F00F00F00Full;
  x = (x | (x << 4)) & 0x30C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x9249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256)
lbvh_keys(const float* __restrict__ leaf_box, unsigned n, float3 lo, float3 inv_extent, unsigned long long* __restrict__ keys,
          unsigned* __restrict__ vals) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* b = leaf_box + 6 * (size_t)i;
    const float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]);
    auto q = [](float c, float l, float inv) {
      float u = (c - l) * inv;
      u = fminf(fmaxf(u, 0.f), 0.9999999f);  // NaN -> 0
      return (unsigned)(u * 65536.f);
    };
    keys[i] = spread16(q(cx, lo.x, inv_extent.x)) | (spread16(q(cy, lo.y, inv_extent.y)) << 1) | (spread16(q(cz, lo.z, inv_extent.z)) << 2);
    vals[i] = i;
  }
}

// ---- radix sort pass -----------------------------------------------------------------------------------
// hist[digit * n_blocks + block] = pairs of that digit in the block's tile
__global__ void __launch_bounds__(kSortBlock)
radix_count(const unsigned long long* __restrict__ keys, unsigned n, int shift, unsigned* __restrict__ hist, unsigned n_blocks) {
  __shared__ unsigned s_cnt[kRadix];
  if (threadIdx.x < kRadix) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const size_t first = (size_t)blockIdx.x * kSortTile + (size_t)threadIdx.x * kSortItems;
  unsigned local[kRadix];
#pragma unroll
  for (int d = 0; d < kRadix; ++d) local[d] = 0u;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const size_t i = first + k;
    if (i < n) {
      const unsigned d = (unsigned)(keys[i] >> shift) & (kRadix - 1);
#pragma unroll
      for (int e = 0; e < kRadix; ++e) local[e] += (d == (unsigned)e);
    }
  }
#pragma unroll
  for (int d = 0; d < kRadix; ++d)
    if (local[d]) atomicAdd(&s_cnt[d], local[d]);
  __syncthreads();
  if (threadIdx.x < kRadix) hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of the whole [digit][block] table by one block (n_entries = 16 * n_blocks <= a few thousand)
__global__ void __launch_bounds__(1024)
radix_scan(unsigned* __restrict__ hist, unsigned n_entries) {
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0u;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < n_entries; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < n_entries ? hist[i] : 0u;
    unsigned incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_warp[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (lane >= (unsigned)o) w += y;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned before = s_carry + (warp ? s_warp[warp - 1] : 0u);
    if (i < n_entries) hist[i] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = before + incl;
    __syncthreads();
  }
}

// ordered scatter: thread t's pairs follow those of threads < t of the same digit (per-digit block scan), and a
// thread writes its own pairs in order
__global__ void __launch_bounds__(kSortBlock)
radix_scatter(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, unsigned n, int shift,
              const unsigned* __restrict__ hist, unsigned n_blocks, unsigned long long* __restrict__ keys_out,
              unsigned* __restrict__ vals_out) {
  __shared__ unsigned s_warp[kRadix][kSortBlock / 32];
  const size_t first = (size_t)blockIdx.x * kSortTile + (size_t)threadIdx.x * kSortItems;
  unsigned long long my_key[kSortItems];
  unsigned my_digit[kSortItems];
  unsigned local[kRadix];
#pragma unroll
  for (int d = 0; d < kRadix; ++d) local[d] = 0u;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const size_t i = first + k;
    my_key[k] = i < n ? keys[i] : 0ull;
    my_digit[k] = i < n ? ((unsigned)(my_key[k] >> shift) & (kRadix - 1)) : 0xFFu;
#pragma unroll
    for (int e = 0; e < kRadix; ++e) local[e] += (my_digit[k] == (unsigned)e);
  }
  // exclusive scan of local[d] over the block's threads, for every digit
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned offs[kRadix];
#pragma unroll
  for (int d = 0; d < kRadix; ++d) {
    unsigned incl = local[d];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31u) s_warp[d][warp] = incl;
    offs[d] = incl - local[d];
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < kRadix; ++d) {
    unsigned before = 0u;
    for (unsigned w = 0; w < warp; ++w) before += s_warp[d][w];
    offs[d] += before + hist[(size_t)d * n_blocks + blockIdx.x];
  }
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const size_t i = first + k;
    if (i < n) {
      unsigned dst = 0u;
#pragma unroll
      for (int e = 0; e < kRadix; ++e)
        if (my_digit[k] == (unsigned)e) dst = offs[e]++;
      keys_out[dst] = my_key[k];
      vals_out[dst] = vals[i];
    }
  }
}

// ---- Karras hierarchy ------------------------------------------------------------------------------------
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

// child / parent encoding while building: inner node k -> k, sorted leaf p -> ~p
__global__ void __launch_bounds__(256)
lbvh_hierarchy(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children, int* __restrict__ parent_inner,
               int* __restrict__ parent_leaf) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
      if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
      if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
      if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? ~gamma : gamma;
    const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    if (left >= 0) parent_inner[left] = i; else parent_leaf[~left] = i;
    if (right >= 0) parent_inner[right] = i; else parent_leaf[~right] = i;
    if (i == 0) parent_inner[0] = -1;
  }
}

struct NodeOut {  // = tutu_internal.hpp InnerNode (64 bytes)
  float box[12];
  int left, right, pad0, pad1;
};

__device__ __forceinline__ void store_child_box(NodeOut* node, int right, const float lo[3], const float hi[3]) {
  float* b = node->box + 6 * right;
  b[0] = lo[0], b[1] = hi[0], b[2] = lo[1], b[3] = hi[1], b[4] = lo[2], b[5] = hi[2];
}

// one thread per sorted leaf; the second thread to reach a node continues with the union of both child boxes
__global__ void __launch_bounds__(256)
lbvh_refit(const float* __restrict__ leaf_box, const unsigned* __restrict__ leaf_code, const unsigned* __restrict__ sorted_leaf,
           int n, const int2* __restrict__ children, const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf,
           unsigned* __restrict__ arrivals, NodeOut* __restrict__ nodes) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const unsigned leaf = sorted_leaf[p];
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) lo[a] = leaf_box[6 * (size_t)leaf + a], hi[a] = leaf_box[6 * (size_t)leaf + 3 + a];
    int child = ~p;
    int node = parent_leaf[p];
    while (node >= 0) {
      const int2 ch = children[node];
      const int right = ch.y == child ? 1 : 0;
      NodeOut* out = nodes + node;
      store_child_box(out, right, lo, hi);
      const int ref = child >= 0 ? child : (int)~leaf_code[sorted_leaf[~child]];
      if (right) out->right = ref; else out->left = ref;
      __threadfence();
      if (atomicAdd(arrivals + node, 1u) == 0u) break;  // the sibling subtree is not finished yet
      __threadfence();
      const volatile float* sib = out->box + 6 * (1 - right);
      for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(lo[a], sib[2 * a]);
        hi[a] = fmaxf(hi[a], sib[2 * a + 1]);
      }
      child = node;
      node = parent_inner[node];
    }
  }
}

__global__ void __launch_bounds__(256)
lbvh_depth(int n, const int* __restrict__ parent_inner, const int* __restrict__ parent_leaf, unsigned* __restrict__ deepest) {
  unsigned best = 0u;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    unsigned d = 0u;
    for (int node = parent_leaf[p]; node >= 0; node = parent_inner[node]) ++d;
    best = max(best, d);
  }
  for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_down_sync(0xFFFFFFFFu, best, o));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(deepest, best);
}


// ---- PLOC: parallel locally-ordered clustering ---------------------------------------------------------------
// Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction" (TVCG 2018).
// Clusters start as the Morton-sorted leaves.  One iteration: every cluster finds, among the kPlocRadius clusters on
// either side of it in the array, the one whose union with it has the smallest surface area (ties: the lower index);
// pairs that chose each other become one inner node, which takes the place of the pair's first cluster; the array
// is compacted in order (so it stays Morton-ordered) and the next iteration runs on the shorter array.  The distance
// is symmetric and the tie rule total, so every iteration merges at least the globally best pair.  Bottom-up merging
// by surface area gives a tree of SAH quality (measured below) in about thirty passes over a shrinking array.
// Nodes are numbered from n - 2 downwards in creation order, so the last merge is node 0 = the root.
#ifndef TUTU_PLOC_RADIUS
#define TUTU_PLOC_RADIUS 8
#endif
constexpr int kPlocRadius = TUTU_PLOC_RADIUS;
constexpr int kScanBlock = 256, kScanItems = 8, kScanTile = kScanBlock * kScanItems;

struct PlocCtl {
  unsigned n_cur;      // clusters in the current array
  unsigned next_node;  // nodes [next_node, n - 1) are taken
  unsigned merged, kept;  // totals of the current iteration (written by ploc_scan_partials)
};

__device__ __forceinline__ float ploc_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
  const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y),
              dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
  return dx * dy + dy * dz + dz * dx;
}

// cluster p = sorted leaf p: lo.w = bits(~p) (the id: inner node >= 0, sorted leaf ~p)
__global__ void __launch_bounds__(256)
ploc_init(const float* __restrict__ leaf_box, const unsigned* __restrict__ sorted_leaf, unsigned n, float4* __restrict__ lo,
          float4* __restrict__ hi, PlocCtl* ctl) {
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const float* b = leaf_box + 6 * (size_t)sorted_leaf[p];
    lo[p] = make_float4(b[0], b[1], b[2], __int_as_float(~(int)p));
    hi[p] = make_float4(b[3], b[4], b[5], 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_cur = n, ctl->next_node = n - 1, ctl->merged = 0, ctl->kept = 0;
}

__global__ void __launch_bounds__(256)
ploc_nearest(const float4* __restrict__ lo, const float4* __restrict__ hi, const PlocCtl* __restrict__ ctl, unsigned* __restrict__ nn) {
  const unsigned n = ctl->n_cur;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 alo = lo[i], ahi = hi[i];
    const unsigned first = i > (unsigned)kPlocRadius ? i - kPlocRadius : 0u, last = min(n - 1u, i + kPlocRadius);
    float best = FLT_MAX;
    unsigned arg = i;
    for (unsigned j = first; j <= last; ++j) {
      if (j == i) continue;
      const float a = ploc_area(alo, ahi, lo[j], hi[j]);
      if (a < best) best = a, arg = j;  // strict: the lower index wins a tie
    }
    nn[i] = arg;
  }
}

// flags of cluster i: merged = it is the first of a mutual pair (becomes a node), kept = it survives the iteration
__device__ __forceinline__ unsigned ploc_flags(const unsigned* __restrict__ nn, unsigned i) {
  const unsigned j = nn[i];
  const bool mutual = j != i && nn[j] == i;
  const unsigned merged = mutual && i < j, kept = !(mutual && i > j);
  return merged | (kept << 16);
}

// per-tile sums of (merged, kept), packed 2 x 16 bits per thread, 2 x 32 per tile
__global__ void __launch_bounds__(kScanBlock)
ploc_count(const unsigned* __restrict__ nn, const PlocCtl* __restrict__ ctl, uint2* __restrict__ tile_sum) {
  const unsigned n = ctl->n_cur;
  __shared__ unsigned s_m[kScanBlock / 32], s_k[kScanBlock / 32];
  for (unsigned tile = blockIdx.x; (size_t)tile * kScanTile < n; tile += gridDim.x) {
    const unsigned first = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned acc = 0u;
    for (int k = 0; k < kScanItems; ++k)
      if (first + k < n) acc += ploc_flags(nn, first + k);
    unsigned m = acc & 0xFFFFu, kp = acc >> 16;
    for (int o = 16; o > 0; o >>= 1) m += __shfl_down_sync(0xFFFFFFFFu, m, o), kp += __shfl_down_sync(0xFFFFFFFFu, kp, o);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m, s_k[threadIdx.x >> 5] = kp;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tm = 0, tk = 0;
      for (int w = 0; w < kScanBlock / 32; ++w) tm += s_m[w], tk += s_k[w];
      tile_sum[tile] = make_uint2(tm, tk);
    }
    __syncthreads();
  }
}

// exclusive scan of the tile sums by one block; totals into ctl
__global__ void __launch_bounds__(1024)
ploc_scan_tiles(uint2* __restrict__ tile_sum, PlocCtl* ctl) {
  const unsigned n_tiles = (ctl->n_cur + kScanTile - 1) / kScanTile;
  __shared__ unsigned s_m[1024], s_k[1024];
  __shared__ unsigned carry_m, carry_k;
  if (threadIdx.x == 0) carry_m = 0, carry_k = 0;
  __syncthreads();
  for (unsigned base = 0; base < n_tiles; base += 1024) {
    const unsigned t = base + threadIdx.x;
    const uint2 v = t < n_tiles ? tile_sum[t] : make_uint2(0, 0);
    s_m[threadIdx.x] = v.x, s_k[threadIdx.x] = v.y;
    __syncthreads();
    for (unsigned o = 1; o < 1024; o <<= 1) {
      const unsigned am = threadIdx.x >= o ? s_m[threadIdx.x - o] : 0u, ak = threadIdx.x >= o ? s_k[threadIdx.x - o] : 0u;
      __syncthreads();
      s_m[threadIdx.x] += am, s_k[threadIdx.x] += ak;
      __syncthreads();
    }
    if (t < n_tiles) tile_sum[t] = make_uint2(carry_m + s_m[threadIdx.x] - v.x, carry_k + s_k[threadIdx.x] - v.y);
    __syncthreads();
    if (threadIdx.x == 1023) carry_m += s_m[1023], carry_k += s_k[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) ctl->merged = carry_m, ctl->kept = carry_k;
}

__device__ __forceinline__ int ploc_final_ref(int id, const unsigned* __restrict__ leaf_code, const unsigned* __restrict__ sorted_leaf) {
  return id >= 0 ? id : (int)~leaf_code[sorted_leaf[~id]];
}

// writes the nodes of the mutual pairs and the next cluster array (order preserved)
__global__ void __launch_bounds__(kScanBlock)
ploc_merge(const float4* __restrict__ lo, const float4* __restrict__ hi, const unsigned* __restrict__ nn, const uint2* __restrict__ tile_sum,
           const PlocCtl* __restrict__ ctl, const unsigned* __restrict__ leaf_code, const unsigned* __restrict__ sorted_leaf,
           float4* __restrict__ lo_out, float4* __restrict__ hi_out, NodeOut* __restrict__ nodes, int* __restrict__ parent_inner,
           int* __restrict__ parent_leaf) {
  const unsigned n = ctl->n_cur, next_node = ctl->next_node;
  __shared__ unsigned s_m[kScanBlock], s_k[kScanBlock];
  for (unsigned tile = blockIdx.x; (size_t)tile * kScanTile < n; tile += gridDim.x) {
    const unsigned first = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned fl[kScanItems];
    unsigned acc = 0u;
    for (int k = 0; k < kScanItems; ++k) {
      fl[k] = first + k < n ? ploc_flags(nn, first + k) : 0u;
      acc += fl[k];
    }
    s_m[threadIdx.x] = acc & 0xFFFFu, s_k[threadIdx.x] = acc >> 16;
    __syncthreads();
    for (unsigned o = 1; o < kScanBlock; o <<= 1) {
      const unsigned am = threadIdx.x >= o ? s_m[threadIdx.x - o] : 0u, ak = threadIdx.x >= o ? s_k[threadIdx.x - o] : 0u;
      __syncthreads();
      s_m[threadIdx.x] += am, s_k[threadIdx.x] += ak;
      __syncthreads();
    }
    const uint2 base = tile_sum[tile];
    unsigned rank_m = base.x + s_m[threadIdx.x] - (acc & 0xFFFFu), rank_k = base.y + s_k[threadIdx.x] - (acc >> 16);
    __syncthreads();
    for (int k = 0; k < kScanItems; ++k) {
      const unsigned i = first + k;
      if (i >= n) break;
      const bool merged = fl[k] & 1u, kept = (fl[k] >> 16) & 1u;
      if (merged) {
        const unsigned j = nn[i];
        const int node = (int)(next_node - 1u - rank_m);
        const float4 alo = lo[i], ahi = hi[i], blo = lo[j], bhi = hi[j];
        const int ida = __float_as_int(alo.w), idb = __float_as_int(blo.w);
        NodeOut* out = nodes + node;
        const float l0[3] = {alo.x, alo.y, alo.z}, h0[3] = {ahi.x, ahi.y, ahi.z}, l1[3] = {blo.x, blo.y, blo.z}, h1[3] = {bhi.x, bhi.y, bhi.z};
        store_child_box(out, 0, l0, h0);
        store_child_box(out, 1, l1, h1);
        out->left = ploc_final_ref(ida, leaf_code, sorted_leaf);
        out->right = ploc_final_ref(idb, leaf_code, sorted_leaf);
        out->pad0 = out->pad1 = 0;
        if (ida >= 0) parent_inner[ida] = node; else parent_leaf[~ida] = node;
        if (idb >= 0) parent_inner[idb] = node; else parent_leaf[~idb] = node;
        lo_out[rank_k] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), __int_as_float(node));
        hi_out[rank_k] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
      } else if (kept) {
        lo_out[rank_k] = lo[i];
        hi_out[rank_k] = hi[i];
      }
      rank_m += merged, rank_k += kept;
    }
  }
}

__global__ void ploc_advance(PlocCtl* ctl, int* __restrict__ parent_inner) {
  ctl->next_node -= ctl->merged;
  ctl->n_cur = ctl->kept;
  if (ctl->kept == 1u) parent_inner[0] = -1;  // the last merge was the root
}


// ---- top-down binned SAH, the host builder's split rule on the device ------------------------------------------
// host_scene.cpp (FastBuilder) splits a node at the cheapest of 3 x 15 bin boundaries of its centroid bounds,
// cost = area(left) * n_left + area(right) * n_right, first minimum in (axis, boundary) order; a node's split depends
// only on the SET of its primitives, so a level-synchronous device build that repeats the arithmetic operation for
// operation (round-to-nearest intrinsics, no contraction) grows the same tree — node for node, in the same pre-order
// numbering (left child = base + 1, right child = base + n_left) — except where the host falls back to halving a node
// by its current order (depth budget exhausted, or all centroids equal), which depends on the order of an unstable
// partition.  Nodes of more than kSahSmall primitives are split level by level, all nodes of a level at once:
//   sah_cb     centroid bounds (segmented warp reduction -> atomics on order-preserving integer images of the floats)
//   sah_bin    16 bins x 3 axes: counts and exact boxes, privatised per block in shared memory, merged by atomics
//   sah_split  one warp per node: 45 candidates, the split record, the two children (leaf / small / large)
//   sah_flags* + sah_partition   stable partition of every node's records by an exclusive scan of the "left" flags
// and every node of at most kSahSmall primitives is finished by one warp in shared memory (sah_small).  Boxes are
// not tracked on the way down: lbvh_refit computes all of them bottom-up as exact fmin/fmax unions afterwards.
constexpr int kSahBins = 16, kSahCand = 3 * (kSahBins - 1);
constexpr unsigned kSahSmall = 64;
constexpr int kSahSlots = 5;  // large nodes (> 64 records) that a block of 256 consecutive records can touch
constexpr int kSahAccWords = 6 + 3 * kSahBins * 7;  // per large node: centroid bounds, then per (axis, bin) count + box

struct SahNode {
  unsigned first, count, base, depth;
};
struct SahSplit {
  int axis, split;      // binned: records whose bin on `axis` is < split go left
  unsigned nl;
  float cb_lo, scale;
  int by_position;      // 1: the first nl records go left (no usable split, or the depth budget is short)
  int child[2];         // index in the next level's list when the child is large, else -1
  unsigned first, count;
};
struct SahCtl {
  unsigned n_nodes[2];  // large nodes of the current / next level
  unsigned n_small;
  unsigned pad;
};

__device__ __forceinline__ unsigned sah_enc(float f) {  // order-preserving image: enc(a) < enc(b) <=> a < b
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sah_dec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }
__device__ __forceinline__ float sah_centroid(float lo, float hi) { return __fmul_rn(0.5f, __fadd_rn(lo, hi)); }
__device__ __forceinline__ int sah_bin_of(float c, float cb_lo, float scale) {
  const int b = (int)__fmul_rn(__fsub_rn(c, cb_lo), scale);
  return b < 0 ? 0 : (b >= kSahBins ? kSahBins - 1 : b);
}
__device__ __forceinline__ float sah_area(const float lo[3], const float hi[3]) {
  const float dx = __fsub_rn(hi[0], lo[0]), dy = __fsub_rn(hi[1], lo[1]), dz = __fsub_rn(hi[2], lo[2]);
  return __fmul_rn(2.f, __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx)));
}
__device__ __forceinline__ int sah_ceil_log2(unsigned n) { return n <= 1u ? 0 : 32 - __clz(n - 1u); }

__global__ void __launch_bounds__(256)
sah_init(const float* __restrict__ leaf_box, unsigned n, float4* __restrict__ lo, float4* __restrict__ hi, int* __restrict__ owner,
         SahNode* __restrict__ nodes, SahNode* __restrict__ small, SahCtl* ctl, int* __restrict__ parent_inner) {
  const bool large = n > kSahSmall;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* b = leaf_box + 6 * (size_t)i;
    lo[i] = make_float4(b[0], b[1], b[2], __uint_as_float(i));
    hi[i] = make_float4(b[3], b[4], b[5], 0.f);
    owner[i] = large ? 0 : -1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->n_nodes[0] = large ? 1u : 0u, ctl->n_nodes[1] = 0u, ctl->n_small = large ? 0u : 1u, ctl->pad = 0u;
    (large ? nodes : small)[0] = SahNode{0u, n, 0u, 0u};
    parent_inner[0] = -1;
  }
}

__global__ void __launch_bounds__(256)
sah_reset(const SahCtl* __restrict__ ctl, unsigned* __restrict__ acc) {
  const size_t total = (size_t)ctl->n_nodes[0] * kSahAccWords;
  const unsigned lo_init = sah_enc(FLT_MAX), hi_init = sah_enc(-FLT_MAX);
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
    const unsigned r = (unsigned)(w % kSahAccWords);
    unsigned v;
    if (r < 6u) v = r < 3u ? lo_init : hi_init;
    else {
      const unsigned q = (r - 6u) % 7u;  // 0 count, 1..3 lo, 4..6 hi
      v = q == 0u ? 0u : (q <= 3u ? lo_init : hi_init);
    }
    acc[w] = v;
  }
}

__global__ void __launch_bounds__(256)
sah_cb(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ owner, unsigned n, unsigned* __restrict__ acc) {
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned i = base + threadIdx.x;
    int k = -1;
    float c[3] = {0.f, 0.f, 0.f};
    if (i < n && (k = owner[i]) >= 0) {
      const float4 l = lo[i], h = hi[i];
      c[0] = sah_centroid(l.x, h.x), c[1] = sah_centroid(l.y, h.y), c[2] = sah_centroid(l.z, h.z);
    }
    float mn[3] = {c[0], c[1], c[2]}, mx[3] = {c[0], c[1], c[2]};
    for (unsigned off = 1; off < 32u; off <<= 1) {  // segmented reduction: a node's records are contiguous
      const int ok = __shfl_down_sync(0xFFFFFFFFu, k, off);
      const bool take = lane + off < 32u && ok == k;
      for (int a = 0; a < 3; ++a) {
        const float vn = __shfl_down_sync(0xFFFFFFFFu, mn[a], off), vx = __shfl_down_sync(0xFFFFFFFFu, mx[a], off);
        if (take) mn[a] = fminf(mn[a], vn), mx[a] = fmaxf(mx[a], vx);
      }
    }
    const int kp = __shfl_up_sync(0xFFFFFFFFu, k, 1);
    if (k >= 0 && (lane == 0u || kp != k)) {
      unsigned* a = acc + (size_t)k * kSahAccWords;
      for (int d = 0; d < 3; ++d) atomicMin(a + d, sah_enc(mn[d])), atomicMax(a + 3 + d, sah_enc(mx[d]));
    }
  }
}

__global__ void __launch_bounds__(256)
sah_bin(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ owner, unsigned n, unsigned* __restrict__ acc) {
  __shared__ unsigned s_acc[kSahSlots][3 * kSahBins * 7];
  __shared__ int s_node[kSahSlots];
  __shared__ unsigned s_warp[8];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lo_init = sah_enc(FLT_MAX), hi_init = sah_enc(-FLT_MAX);
  for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    for (unsigned w = threadIdx.x; w < kSahSlots * 3 * kSahBins * 7; w += blockDim.x) {
      const unsigned q = w % 7u;
      (&s_acc[0][0])[w] = q == 0u ? 0u : (q <= 3u ? lo_init : hi_init);
    }
    if (threadIdx.x < kSahSlots) s_node[threadIdx.x] = -1;
    const unsigned i = base + threadIdx.x;
    const int k = i < n ? owner[i] : -1;
    const int kp = (i < n && i > 0u && threadIdx.x > 0u) ? owner[i - 1] : -2;
    const bool start = k >= 0 && kp != k;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, start);
    if (lane == 0u) s_warp[warp] = (unsigned)__popc(bal);
    __syncthreads();
    unsigned slot = (unsigned)__popc(bal & ((2u << lane) - 1u));  // starts up to and including this lane
    for (unsigned w = 0; w < warp; ++w) slot += s_warp[w];
    // slot - 1 = index of the run this record belongs to (runs of records without an owner do not count)
    if (start && slot <= (unsigned)kSahSlots) s_node[slot - 1u] = k;
    __syncthreads();
    if (k >= 0) {
      const unsigned* cb = acc + (size_t)k * kSahAccWords;
      const float4 l = lo[i], h = hi[i];
      const float bl[3] = {l.x, l.y, l.z}, bh[3] = {h.x, h.y, h.z};
      const bool priv = slot >= 1u && slot <= (unsigned)kSahSlots;
      unsigned* dst = priv ? &s_acc[slot - 1u][0] : acc + (size_t)k * kSahAccWords + 6;
      for (int a = 0; a < 3; ++a) {
        const float clo = sah_dec(cb[a]), chi = sah_dec(cb[3 + a]);
        const float ext = __fsub_rn(chi, clo);
        if (!(ext > 0.f)) continue;
        const int b = sah_bin_of(sah_centroid(bl[a], bh[a]), clo, __fdiv_rn((float)kSahBins, ext));
        unsigned* e = dst + (a * kSahBins + b) * 7;
        atomicAdd(e, 1u);
        for (int d = 0; d < 3; ++d) atomicMin(e + 1 + d, sah_enc(bl[d])), atomicMax(e + 4 + d, sah_enc(bh[d]));
      }
    }
    __syncthreads();
    for (unsigned w = threadIdx.x; w < kSahSlots * 3 * kSahBins; w += blockDim.x) {  // one (slot, axis, bin) entry per thread
      const unsigned sl = w / (3 * kSahBins), e = w % (3 * kSahBins);
      const int node = s_node[sl];
      const unsigned* src = &s_acc[sl][e * 7];
      if (node < 0 || src[0] == 0u) continue;
      unsigned* dst = acc + (size_t)node * kSahAccWords + 6 + e * 7;
      atomicAdd(dst, src[0]);
      for (int d = 0; d < 3; ++d) atomicMin(dst + 1 + d, src[1 + d]), atomicMax(dst + 4 + d, src[4 + d]);
    }
    __syncthreads();
  }
}

// a child of `count` records starting at record `first`, inner-node index `cbase`: leaf, small subtree, or large node
__device__ __forceinline__ int sah_child(unsigned first, unsigned count, unsigned cbase, unsigned depth, unsigned parent, SahNode* next,
                                         SahNode* small, SahCtl* ctl, int* parent_inner, int* parent_leaf, int* large_index) {
  *large_index = -1;
  if (count == 1u) {
    parent_leaf[first] = (int)parent;
    return ~(int)first;
  }
  parent_inner[cbase] = (int)parent;
  if (count <= kSahSmall) {
    small[atomicAdd(&ctl->n_small, 1u)] = SahNode{first, count, cbase, depth};
  } else {
    const unsigned at = atomicAdd(&ctl->n_nodes[1], 1u);
    next[at] = SahNode{first, count, cbase, depth};
    *large_index = (int)at;
  }
  return (int)cbase;
}

__global__ void __launch_bounds__(128)
sah_split(const SahNode* __restrict__ nodes, const unsigned* __restrict__ acc, SahCtl* ctl, SahNode* __restrict__ next, SahNode* __restrict__ small,
          SahSplit* __restrict__ split, int2* __restrict__ children, int* __restrict__ parent_inner, int* __restrict__ parent_leaf, int max_depth) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned m = ctl->n_nodes[0];
  for (unsigned k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < m; k += gridDim.x * (blockDim.x >> 5)) {
    const SahNode nd = nodes[k];
    const unsigned* a0 = acc + (size_t)k * kSahAccWords;
    const bool median = (max_depth - (int)nd.depth) <= sah_ceil_log2(nd.count) + 1;
    float best = FLT_MAX;
    int best_c = 1 << 30;
    if (!median) {
      for (int c = (int)lane; c < kSahCand; c += 32) {
        const int a = c / (kSahBins - 1), b = c % (kSahBins - 1) + 1;  // bins [0, b) left, [b, 16) right
        if (!(__fsub_rn(sah_dec(a0[3 + a]), sah_dec(a0[a])) > 0.f)) continue;
        const unsigned* e = a0 + 6 + a * kSahBins * 7;
        unsigned nl = 0u, nr = 0u;
        unsigned llo[3], lhi[3], rlo[3], rhi[3];
        for (int d = 0; d < 3; ++d) llo[d] = rlo[d] = 0xFFFFFFFFu, lhi[d] = rhi[d] = 0u;
        for (int q = 0; q < kSahBins; ++q) {
          const unsigned* bq = e + q * 7;
          if (q < b) {
            nl += bq[0];
            for (int d = 0; d < 3; ++d) llo[d] = min(llo[d], bq[1 + d]), lhi[d] = max(lhi[d], bq[4 + d]);
          } else {
            nr += bq[0];
            for (int d = 0; d < 3; ++d) rlo[d] = min(rlo[d], bq[1 + d]), rhi[d] = max(rhi[d], bq[4 + d]);
          }
        }
        if (nl == 0u || nr == 0u) continue;
        float fl[3], fh[3], gl[3], gh[3];
        for (int d = 0; d < 3; ++d) fl[d] = sah_dec(llo[d]), fh[d] = sah_dec(lhi[d]), gl[d] = sah_dec(rlo[d]), gh[d] = sah_dec(rhi[d]);
        const float cost = __fadd_rn(__fmul_rn(sah_area(fl, fh), __uint2float_rn(nl)), __fmul_rn(sah_area(gl, gh), __uint2float_rn(nr)));
        if (cost < best) best = cost, best_c = c;
      }
      for (int off = 16; off > 0; off >>= 1) {  // first minimum in candidate order
        const float oc = __shfl_xor_sync(0xFFFFFFFFu, best, off);
        const int oi = __shfl_xor_sync(0xFFFFFFFFu, best_c, off);
        if (oi < (1 << 30) && (best_c == (1 << 30) || oc < best || (oc == best && oi < best_c))) best = oc, best_c = oi;
      }
    }
    if (lane == 0u) {
      SahSplit sp{};
      sp.first = nd.first, sp.count = nd.count;
      if (best_c < kSahCand) {
        sp.axis = best_c / (kSahBins - 1), sp.split = best_c % (kSahBins - 1) + 1, sp.by_position = 0;
        const float clo = sah_dec(a0[sp.axis]);
        sp.cb_lo = clo, sp.scale = __fdiv_rn((float)kSahBins, __fsub_rn(sah_dec(a0[3 + sp.axis]), clo));
        unsigned nl = 0u;
        for (int q = 0; q < sp.split; ++q) nl += a0[6 + (sp.axis * kSahBins + q) * 7];
        sp.nl = nl;
      } else {
        sp.by_position = 1, sp.nl = nd.count / 2u;
      }
      const int lref = sah_child(nd.first, sp.nl, nd.base + 1u, nd.depth + 1u, nd.base, next, small, ctl, parent_inner, parent_leaf, &sp.child[0]);
      const int rref = sah_child(nd.first + sp.nl, nd.count - sp.nl, nd.base + sp.nl, nd.depth + 1u, nd.base, next, small, ctl, parent_inner,
                                 parent_leaf, &sp.child[1]);
      children[nd.base] = make_int2(lref, rref);
      split[k] = sp;
    }
  }
}

__device__ __forceinline__ unsigned sah_goes_left(const float4 l, const float4 h, const SahSplit& sp) {
  const float c = sp.axis == 0 ? sah_centroid(l.x, h.x) : sp.axis == 1 ? sah_centroid(l.y, h.y) : sah_centroid(l.z, h.z);
  return sah_bin_of(c, sp.cb_lo, sp.scale) < sp.split ? 1u : 0u;
}
__device__ __forceinline__ unsigned sah_flag(const float4* lo, const float4* hi, const int* owner, const SahSplit* split, unsigned i) {
  const int k = owner[i];
  if (k < 0) return 0u;
  const SahSplit& sp = split[k];
  return sp.by_position ? 0u : sah_goes_left(lo[i], hi[i], sp);
}

__global__ void __launch_bounds__(kScanBlock)
sah_flag_count(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ owner, const SahSplit* __restrict__ split, unsigned n,
               unsigned* __restrict__ tile_sum) {
  __shared__ unsigned s_w[kScanBlock / 32];
  for (unsigned tile = blockIdx.x; (size_t)tile * kScanTile < n; tile += gridDim.x) {
    const unsigned first = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned acc = 0u;
    for (int q = 0; q < kScanItems; ++q)
      if (first + q < n) acc += sah_flag(lo, hi, owner, split, first + q);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned t = 0;
      for (int w = 0; w < kScanBlock / 32; ++w) t += s_w[w];
      tile_sum[tile] = t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024)
sah_scan_tiles(unsigned* __restrict__ tile_sum, unsigned n) {
  const unsigned n_tiles = (n + kScanTile - 1) / kScanTile;
  __shared__ unsigned s_v[1024];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned base = 0; base < n_tiles; base += 1024) {
    const unsigned t = base + threadIdx.x;
    const unsigned v = t < n_tiles ? tile_sum[t] : 0u;
    s_v[threadIdx.x] = v;
    __syncthreads();
    for (unsigned o = 1; o < 1024; o <<= 1) {
      const unsigned add = threadIdx.x >= o ? s_v[threadIdx.x - o] : 0u;
      __syncthreads();
      s_v[threadIdx.x] += add;
      __syncthreads();
    }
    if (t < n_tiles) tile_sum[t] = carry + s_v[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s_v[1023];
    __syncthreads();
  }
}

// S[i] = number of "left" flags before record i (exclusive scan over the whole array)
__global__ void __launch_bounds__(kScanBlock)
sah_flag_scan(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ owner, const SahSplit* __restrict__ split, unsigned n,
              const unsigned* __restrict__ tile_sum, unsigned* __restrict__ S) {
  __shared__ unsigned s_v[kScanBlock];
  for (unsigned tile = blockIdx.x; (size_t)tile * kScanTile < n; tile += gridDim.x) {
    const unsigned first = tile * kScanTile + threadIdx.x * kScanItems;
    unsigned fl[kScanItems], acc = 0u;
    for (int q = 0; q < kScanItems; ++q) {
      fl[q] = first + q < n ? sah_flag(lo, hi, owner, split, first + q) : 0u;
      acc += fl[q];
    }
    s_v[threadIdx.x] = acc;
    __syncthreads();
    for (unsigned o = 1; o < kScanBlock; o <<= 1) {
      const unsigned add = threadIdx.x >= o ? s_v[threadIdx.x - o] : 0u;
      __syncthreads();
      s_v[threadIdx.x] += add;
      __syncthreads();
    }
    unsigned run = tile_sum[tile] + s_v[threadIdx.x] - acc;
    __syncthreads();
    for (int q = 0; q < kScanItems; ++q) {
      if (first + q >= n) break;
      S[first + q] = run;
      run += fl[q];
    }
  }
}

__global__ void __launch_bounds__(256)
sah_partition(const float4* __restrict__ lo, const float4* __restrict__ hi, const int* __restrict__ owner, const SahSplit* __restrict__ split,
              const unsigned* __restrict__ S, unsigned n, float4* __restrict__ lo_out, float4* __restrict__ hi_out, int* __restrict__ owner_out) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 l = lo[i], h = hi[i];
    const int k = owner[i];
    unsigned dest = i;
    int next_owner = -1;
    if (k >= 0) {
      const SahSplit sp = split[k];
      unsigned left;
      if (sp.by_position) {
        left = (i - sp.first) < sp.nl ? 1u : 0u;
      } else {
        left = sah_goes_left(l, h, sp);
        const unsigned lefts_before = S[i] - S[sp.first];
        dest = left ? sp.first + lefts_before : sp.first + sp.nl + ((i - sp.first) - lefts_before);
      }
      next_owner = sp.child[left ? 0 : 1];
    }
    lo_out[dest] = l, hi_out[dest] = h, owner_out[dest] = next_owner;
  }
}

__global__ void sah_advance(SahCtl* ctl) {
  ctl->n_nodes[0] = ctl->n_nodes[1];
  ctl->n_nodes[1] = 0u;
}

// one warp finishes one subtree of at most kSahSmall records in shared memory (same split rule, candidates evaluated
// directly from the records: the union over "bin < b" is what the host's bin sweep accumulates)
constexpr int kSahSmallWarps = 4;
__global__ void __launch_bounds__(kSahSmallWarps * 32)
sah_small(const SahNode* __restrict__ small, const SahCtl* __restrict__ ctl, float4* __restrict__ lo, float4* __restrict__ hi, int2* __restrict__ children,
          int* __restrict__ parent_inner, int* __restrict__ parent_leaf, int max_depth) {
  __shared__ float4 s_lo[kSahSmallWarps][kSahSmall], s_hi[kSahSmallWarps][kSahSmall];
  __shared__ unsigned char s_bin[kSahSmallWarps][kSahSmall][4];
  __shared__ uint4 s_stack[kSahSmallWarps][kSahSmall];
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  float4* L = s_lo[w];
  float4* H = s_hi[w];
  const unsigned lt = (1u << lane) - 1u;
  const unsigned total = ctl->n_small;
  for (unsigned job = blockIdx.x * kSahSmallWarps + w; job < total; job += gridDim.x * kSahSmallWarps) {
    const SahNode root = small[job];
    for (unsigned r = lane; r < root.count; r += 32u) L[r] = lo[root.first + r], H[r] = hi[root.first + r];
    int sp = 0;
    if (lane == 0u) s_stack[w][0] = make_uint4(0u, root.count, root.base, root.depth);
    sp = 1;
    __syncwarp();
    while (sp > 0) {
      const uint4 nd = s_stack[w][--sp];
      __syncwarp();
      const unsigned f = nd.x, c = nd.y, base = nd.z, depth = nd.w;
      unsigned nl = c / 2u;
      const bool median = (max_depth - (int)depth) <= sah_ceil_log2(c) + 1;
      if (!median) {
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (unsigned r = lane; r < c; r += 32u) {
          const float4 l = L[f + r], h = H[f + r];
          const float cc[3] = {sah_centroid(l.x, h.x), sah_centroid(l.y, h.y), sah_centroid(l.z, h.z)};
          for (int a = 0; a < 3; ++a) mn[a] = fminf(mn[a], cc[a]), mx[a] = fmaxf(mx[a], cc[a]);
        }
        for (int off = 16; off > 0; off >>= 1)
          for (int a = 0; a < 3; ++a)
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xFFFFFFFFu, mn[a], off)), mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xFFFFFFFFu, mx[a], off));
        bool use[3];
        float scale[3];
        for (int a = 0; a < 3; ++a) {
          const float ext = __fsub_rn(mx[a], mn[a]);
          use[a] = ext > 0.f;
          scale[a] = use[a] ? __fdiv_rn((float)kSahBins, ext) : 0.f;
        }
        for (unsigned r = lane; r < c; r += 32u) {
          const float4 l = L[f + r], h = H[f + r];
          s_bin[w][f + r][0] = (unsigned char)sah_bin_of(sah_centroid(l.x, h.x), mn[0], scale[0]);
          s_bin[w][f + r][1] = (unsigned char)sah_bin_of(sah_centroid(l.y, h.y), mn[1], scale[1]);
          s_bin[w][f + r][2] = (unsigned char)sah_bin_of(sah_centroid(l.z, h.z), mn[2], scale[2]);
        }
        __syncwarp();
        float best = FLT_MAX;
        int best_c = 1 << 30;
        for (int cand = (int)lane; cand < kSahCand; cand += 32) {
          const int a = cand / (kSahBins - 1), b = cand % (kSahBins - 1) + 1;
          if (!use[a]) continue;
          float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
          float rlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
          unsigned cl = 0u, cr = 0u;
          for (unsigned r = 0; r < c; ++r) {
            const float4 l = L[f + r], h = H[f + r];
            if ((int)s_bin[w][f + r][a] < b) {
              ++cl;
              llo[0] = fminf(llo[0], l.x), llo[1] = fminf(llo[1], l.y), llo[2] = fminf(llo[2], l.z);
              lhi[0] = fmaxf(lhi[0], h.x), lhi[1] = fmaxf(lhi[1], h.y), lhi[2] = fmaxf(lhi[2], h.z);
            } else {
              ++cr;
              rlo[0] = fminf(rlo[0], l.x), rlo[1] = fminf(rlo[1], l.y), rlo[2] = fminf(rlo[2], l.z);
              rhi[0] = fmaxf(rhi[0], h.x), rhi[1] = fmaxf(rhi[1], h.y), rhi[2] = fmaxf(rhi[2], h.z);
            }
          }
          if (cl == 0u || cr == 0u) continue;
          const float cost = __fadd_rn(__fmul_rn(sah_area(llo, lhi), __uint2float_rn(cl)), __fmul_rn(sah_area(rlo, rhi), __uint2float_rn(cr)));
          if (cost < best) best = cost, best_c = cand;
        }
        for (int off = 16; off > 0; off >>= 1) {
          const float oc = __shfl_xor_sync(0xFFFFFFFFu, best, off);
          const int oi = __shfl_xor_sync(0xFFFFFFFFu, best_c, off);
          if (oi < (1 << 30) && (best_c == (1 << 30) || oc < best || (oc == best && oi < best_c))) best = oc, best_c = oi;
        }
        if (best_c < kSahCand) {  // stable partition of the node's records
          const int a = best_c / (kSahBins - 1), b = best_c % (kSahBins - 1) + 1;
          const bool v0 = lane < c, v1 = lane + 32u < c;
          const bool l0 = v0 && (int)s_bin[w][f + lane][a] < b, l1 = v1 && (int)s_bin[w][f + lane + 32u][a] < b;
          const unsigned m0 = __ballot_sync(0xFFFFFFFFu, l0), m1 = __ballot_sync(0xFFFFFFFFu, l1);
          nl = (unsigned)(__popc(m0) + __popc(m1));
          const unsigned p0 = l0 ? (unsigned)__popc(m0 & lt) : nl + (lane - (unsigned)__popc(m0 & lt));
          const unsigned p1 = l1 ? (unsigned)(__popc(m0) + __popc(m1 & lt)) : nl + (32u - (unsigned)__popc(m0)) + (lane - (unsigned)__popc(m1 & lt));
          float4 a0 = make_float4(0, 0, 0, 0), b0 = a0, a1 = a0, b1 = a0;
          if (v0) a0 = L[f + lane], b0 = H[f + lane];
          if (v1) a1 = L[f + lane + 32u], b1 = H[f + lane + 32u];
          __syncwarp();
          if (v0) L[f + p0] = a0, H[f + p0] = b0;
          if (v1) L[f + p1] = a1, H[f + p1] = b1;
          __syncwarp();
        }
      }
      if (lane == 0u) {
        const unsigned gl = root.first + f;  // global record index of the node's first record
        int lref, rref;
        if (nl == 1u) lref = ~(int)gl, parent_leaf[gl] = (int)base;
        else {
          lref = (int)(base + 1u), parent_inner[base + 1u] = (int)base;
          s_stack[w][sp] = make_uint4(f, nl, base + 1u, depth + 1u);
        }
        const unsigned nr = c - nl;
        if (nr == 1u) rref = ~(int)(gl + nl), parent_leaf[gl + nl] = (int)base;
        else {
          rref = (int)(base + nl), parent_inner[base + nl] = (int)base;
          s_stack[w][sp + (nl == 1u ? 0 : 1)] = make_uint4(f + nl, nr, base + nl, depth + 1u);
        }
        children[base] = make_int2(lref, rref);
      }
      sp += (nl == 1u ? 0 : 1) + ((c - nl) == 1u ? 0 : 1);
      __syncwarp();
    }
    for (unsigned r = lane; r < root.count; r += 32u) lo[root.first + r] = L[r], hi[root.first + r] = H[r];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
sah_ids(const float4* __restrict__ lo, unsigned n, unsigned* __restrict__ ids) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ids[i] = __float_as_uint(lo[i].w);
}

}  // namespace

size_t device_build_lbvh_scratch_bytes(uint32_t n) {
  const unsigned n_blocks = (n + kSortTile - 1) / kSortTile;
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return 2 * pad((size_t)n * 8) + 2 * pad((size_t)n * 4) + pad((size_t)kRadix * n_blocks * 4) + pad((size_t)n * 8) + 3 * pad((size_t)n * 4) + 256;
}

cudaError_t device_build_lbvh(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, const float root_lo[3],
                              const float root_hi[3], void* d_inner_out, uint32_t* depth_out, int sm_count, void* d_scratch,
                              cudaStream_t s) {
  static_assert(sizeof(NodeOut) == 64, "NodeOut must match InnerNode");
  if (n < 2) return cudaErrorInvalidValue;
  const unsigned n_blocks = (n + kSortTile - 1) / kSortTile;
  // scratch: keys x2, vals x2, hist, children, parents x2, arrivals, deepest
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t o_k0 = take((size_t)n * 8), o_k1 = take((size_t)n * 8), o_v0 = take((size_t)n * 4), o_v1 = take((size_t)n * 4);
  const size_t o_hist = take((size_t)kRadix * n_blocks * 4), o_ch = take((size_t)n * 8), o_pi = take((size_t)n * 4);
  const size_t o_pl = take((size_t)n * 4), o_arr = take((size_t)n * 4), o_deep = take(256);
  cudaError_t e = cudaSuccess;
  if (off > device_build_lbvh_scratch_bytes(n) || !d_scratch) return cudaErrorInvalidValue;
  char* base = static_cast<char*>(d_scratch);
  auto* k0 = reinterpret_cast<unsigned long long*>(base + o_k0);
  auto* k1 = reinterpret_cast<unsigned long long*>(base + o_k1);
  auto* v0 = reinterpret_cast<unsigned*>(base + o_v0);
  auto* v1 = reinterpret_cast<unsigned*>(base + o_v1);
  auto* hist = reinterpret_cast<unsigned*>(base + o_hist);
  auto* children = reinterpret_cast<int2*>(base + o_ch);
  auto* parent_inner = reinterpret_cast<int*>(base + o_pi);
  auto* parent_leaf = reinterpret_cast<int*>(base + o_pl);
  auto* arrivals = reinterpret_cast<unsigned*>(base + o_arr);
  auto* deepest = reinterpret_cast<unsigned*>(base + o_deep);

  const int grid = sm_count * 8;
  float3 lo = make_float3(root_lo[0], root_lo[1], root_lo[2]);
  float3 inv;
  inv.x = root_hi[0] > root_lo[0] ? 1.f / (root_hi[0] - root_lo[0]) : 0.f;
  inv.y = root_hi[1] > root_lo[1] ? 1.f / (root_hi[1] - root_lo[1]) : 0.f;
  inv.z = root_hi[2] > root_lo[2] ? 1.f / (root_hi[2] - root_lo[2]) : 0.f;
  lbvh_keys<<<grid, 256, 0, s>>>(d_leaf_box, n, lo, inv, k0, v0);
  for (int shift = 0; shift < kKeyBits; shift += kRadixBits) {
    radix_count<<<n_blocks, kSortBlock, 0, s>>>(k0, n, shift, hist, n_blocks);
    radix_scan<<<1, 1024, 0, s>>>(hist, kRadix * n_blocks);
    radix_scatter<<<n_blocks, kSortBlock, 0, s>>>(k0, v0, n, shift, hist, n_blocks, k1, v1);
    unsigned long long* tk = k0;
    k0 = k1, k1 = tk;
    unsigned* tv = v0;
    v0 = v1, v1 = tv;
  }
  if ((e = cudaMemsetAsync(arrivals, 0, (size_t)n * 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(deepest, 0, 4, s)) != cudaSuccess) return e;
  lbvh_hierarchy<<<grid, 256, 0, s>>>(k0, (int)n, children, parent_inner, parent_leaf);
  lbvh_refit<<<grid, 256, 0, s>>>(d_leaf_box, d_leaf_code, v0, (int)n, children, parent_inner, parent_leaf, arrivals,
                                  static_cast<NodeOut*>(d_inner_out));
  lbvh_depth<<<grid, 256, 0, s>>>((int)n, parent_inner, parent_leaf, deepest);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  unsigned h_deep = 0;
  if ((e = cudaMemcpyAsync(&h_deep, deepest, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  *depth_out = h_deep;
  return cudaSuccess;
}

size_t device_build_ploc_scratch_bytes(uint32_t n) {
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t n_tiles = ((size_t)n + kScanTile - 1) / kScanTile;
  // the Morton sort's buffers (keys x2, vals x2, hist) + cluster arrays x2 (lo, hi) + nn + tile sums + parents x2 + ctl + deepest
  const unsigned n_blocks = (n + kSortTile - 1) / kSortTile;
  return 2 * pad((size_t)n * 8) + 2 * pad((size_t)n * 4) + pad((size_t)kRadix * n_blocks * 4) + 4 * pad((size_t)n * 16) + pad((size_t)n * 4) +
         pad(n_tiles * 8) + 2 * pad((size_t)n * 4) + 2 * 256;
}

cudaError_t device_build_ploc(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, const float root_lo[3],
                              const float root_hi[3], void* d_inner_out, uint32_t* depth_out, uint32_t* iterations_out, int sm_count,
                              void* d_scratch, cudaStream_t s) {
  if (n < 2 || !d_scratch) return cudaErrorInvalidValue;
  const unsigned n_blocks = (n + kSortTile - 1) / kSortTile;
  const size_t n_tiles = ((size_t)n + kScanTile - 1) / kScanTile;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t o_k0 = take((size_t)n * 8), o_k1 = take((size_t)n * 8), o_v0 = take((size_t)n * 4), o_v1 = take((size_t)n * 4);
  const size_t o_hist = take((size_t)kRadix * n_blocks * 4);
  const size_t o_lo0 = take((size_t)n * 16), o_hi0 = take((size_t)n * 16), o_lo1 = take((size_t)n * 16), o_hi1 = take((size_t)n * 16);
  const size_t o_nn = take((size_t)n * 4), o_tiles = take(n_tiles * 8), o_pi = take((size_t)n * 4), o_pl = take((size_t)n * 4);
  const size_t o_ctl = take(256), o_deep = take(256);
  if (off > device_build_ploc_scratch_bytes(n)) return cudaErrorInvalidValue;
  char* base = static_cast<char*>(d_scratch);
  auto* k0 = reinterpret_cast<unsigned long long*>(base + o_k0);
  auto* k1 = reinterpret_cast<unsigned long long*>(base + o_k1);
  auto* v0 = reinterpret_cast<unsigned*>(base + o_v0);
  auto* v1 = reinterpret_cast<unsigned*>(base + o_v1);
  auto* hist = reinterpret_cast<unsigned*>(base + o_hist);
  float4* lo[2] = {reinterpret_cast<float4*>(base + o_lo0), reinterpret_cast<float4*>(base + o_lo1)};
  float4* hi[2] = {reinterpret_cast<float4*>(base + o_hi0), reinterpret_cast<float4*>(base + o_hi1)};
  auto* nn = reinterpret_cast<unsigned*>(base + o_nn);
  auto* tiles = reinterpret_cast<uint2*>(base + o_tiles);
  auto* parent_inner = reinterpret_cast<int*>(base + o_pi);
  auto* parent_leaf = reinterpret_cast<int*>(base + o_pl);
  auto* ctl = reinterpret_cast<PlocCtl*>(base + o_ctl);
  auto* deepest = reinterpret_cast<unsigned*>(base + o_deep);
  cudaError_t e = cudaSuccess;

  const int grid = sm_count * 8;
  float3 lo3 = make_float3(root_lo[0], root_lo[1], root_lo[2]);
  float3 inv;
  inv.x = root_hi[0] > root_lo[0] ? 1.f / (root_hi[0] - root_lo[0]) : 0.f;
  inv.y = root_hi[1] > root_lo[1] ? 1.f / (root_hi[1] - root_lo[1]) : 0.f;
  inv.z = root_hi[2] > root_lo[2] ? 1.f / (root_hi[2] - root_lo[2]) : 0.f;
  lbvh_keys<<<grid, 256, 0, s>>>(d_leaf_box, n, lo3, inv, k0, v0);
  for (int shift = 0; shift < kKeyBits; shift += kRadixBits) {
    radix_count<<<n_blocks, kSortBlock, 0, s>>>(k0, n, shift, hist, n_blocks);
    radix_scan<<<1, 1024, 0, s>>>(hist, kRadix * n_blocks);
    radix_scatter<<<n_blocks, kSortBlock, 0, s>>>(k0, v0, n, shift, hist, n_blocks, k1, v1);
    unsigned long long* tk = k0;
    k0 = k1, k1 = tk;
    unsigned* tv = v0;
    v0 = v1, v1 = tv;
  }
  if ((e = cudaMemsetAsync(deepest, 0, 4, s)) != cudaSuccess) return e;
  ploc_init<<<grid, 256, 0, s>>>(d_leaf_box, v0, n, lo[0], hi[0], ctl);
  // the host looks at the cluster count every few iterations; every iteration merges at least one pair
  const int tile_grid = (int)std::min<size_t>(n_tiles, (size_t)sm_count * 8);
  unsigned iterations = 0;
  int cur = 0;
  PlocCtl h{};
  h.n_cur = n;
  while (h.n_cur > 1u) {
    for (int k = 0; k < 8; ++k, ++iterations) {
      ploc_nearest<<<grid, 256, 0, s>>>(lo[cur], hi[cur], ctl, nn);
      ploc_count<<<tile_grid, kScanBlock, 0, s>>>(nn, ctl, tiles);
      ploc_scan_tiles<<<1, 1024, 0, s>>>(tiles, ctl);
      ploc_merge<<<tile_grid, kScanBlock, 0, s>>>(lo[cur], hi[cur], nn, tiles, ctl, d_leaf_code, v0, lo[cur ^ 1], hi[cur ^ 1],
                                                  static_cast<NodeOut*>(d_inner_out), parent_inner, parent_leaf);
      ploc_advance<<<1, 1, 0, s>>>(ctl, parent_inner);
      cur ^= 1;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if (iterations > 100000u) return cudaErrorUnknown;
  }
  lbvh_depth<<<grid, 256, 0, s>>>((int)n, parent_inner, parent_leaf, deepest);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  unsigned h_deep = 0;
  if ((e = cudaMemcpyAsync(&h_deep, deepest, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  *depth_out = h_deep;
  if (iterations_out) *iterations_out = iterations;
  return cudaSuccess;
}

size_t device_build_sah_scratch_bytes(uint32_t n) {
  auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t n_tiles = ((size_t)n + kScanTile - 1) / kScanTile;
  const size_t max_large = (size_t)n / (kSahSmall + 1) + 2, max_small = (size_t)n / 2 + 2;
  return 4 * pad((size_t)n * 16) + 2 * pad((size_t)n * 4) + pad((size_t)n * 4) + pad(n_tiles * 4) + 2 * pad(max_large * sizeof(SahNode)) +
         pad(max_small * sizeof(SahNode)) + pad(max_large * sizeof(SahSplit)) + pad(max_large * kSahAccWords * 4) + pad((size_t)n * 8) +
         3 * pad((size_t)n * 4) + pad((size_t)n * 4) + 2 * 256;
}

cudaError_t device_build_sah(const float* d_leaf_box, const uint32_t* d_leaf_code, uint32_t n, int max_depth, void* d_inner_out,
                             uint32_t* depth_out, uint32_t* levels_out, int sm_count, void* d_scratch, cudaStream_t s) {
  if (n < 2 || !d_scratch) return cudaErrorInvalidValue;
  const size_t n_tiles = ((size_t)n + kScanTile - 1) / kScanTile;
  const size_t max_large = (size_t)n / (kSahSmall + 1) + 2, max_small = (size_t)n / 2 + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t o_lo0 = take((size_t)n * 16), o_hi0 = take((size_t)n * 16), o_lo1 = take((size_t)n * 16), o_hi1 = take((size_t)n * 16);
  const size_t o_ow0 = take((size_t)n * 4), o_ow1 = take((size_t)n * 4), o_S = take((size_t)n * 4), o_tiles = take(n_tiles * 4);
  const size_t o_nd0 = take(max_large * sizeof(SahNode)), o_nd1 = take(max_large * sizeof(SahNode)), o_small = take(max_small * sizeof(SahNode));
  const size_t o_split = take(max_large * sizeof(SahSplit)), o_acc = take(max_large * kSahAccWords * 4), o_ch = take((size_t)n * 8);
  const size_t o_pi = take((size_t)n * 4), o_pl = take((size_t)n * 4), o_arr = take((size_t)n * 4), o_ids = take((size_t)n * 4);
  const size_t o_ctl = take(256), o_deep = take(256);
  if (off > device_build_sah_scratch_bytes(n)) return cudaErrorInvalidValue;
  char* base = static_cast<char*>(d_scratch);
  float4* lo[2] = {reinterpret_cast<float4*>(base + o_lo0), reinterpret_cast<float4*>(base + o_lo1)};
  float4* hi[2] = {reinterpret_cast<float4*>(base + o_hi0), reinterpret_cast<float4*>(base + o_hi1)};
  int* owner[2] = {reinterpret_cast<int*>(base + o_ow0), reinterpret_cast<int*>(base + o_ow1)};
  auto* S = reinterpret_cast<unsigned*>(base + o_S);
  auto* tiles = reinterpret_cast<unsigned*>(base + o_tiles);
  SahNode* nodes[2] = {reinterpret_cast<SahNode*>(base + o_nd0), reinterpret_cast<SahNode*>(base + o_nd1)};
  auto* small = reinterpret_cast<SahNode*>(base + o_small);
  auto* split = reinterpret_cast<SahSplit*>(base + o_split);
  auto* acc = reinterpret_cast<unsigned*>(base + o_acc);
  auto* children = reinterpret_cast<int2*>(base + o_ch);
  auto* parent_inner = reinterpret_cast<int*>(base + o_pi);
  auto* parent_leaf = reinterpret_cast<int*>(base + o_pl);
  auto* arrivals = reinterpret_cast<unsigned*>(base + o_arr);
  auto* ids = reinterpret_cast<unsigned*>(base + o_ids);
  auto* ctl = reinterpret_cast<SahCtl*>(base + o_ctl);
  auto* deepest = reinterpret_cast<unsigned*>(base + o_deep);
  cudaError_t e = cudaSuccess;
  const int grid = sm_count * 8;
  const int tile_grid = (int)std::min<size_t>(n_tiles, (size_t)sm_count * 8);
  if ((e = cudaMemsetAsync(arrivals, 0, (size_t)n * 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(deepest, 0, 4, s)) != cudaSuccess) return e;
  sah_init<<<grid, 256, 0, s>>>(d_leaf_box, n, lo[0], hi[0], owner[0], nodes[0], small, ctl, parent_inner);
  int cur = 0;
  unsigned levels = 0;
  SahCtl h{};
  h.n_nodes[0] = n > kSahSmall ? 1u : 0u;
  while (h.n_nodes[0] > 0u) {  // one level of large nodes per pass; the host reads the next level's node count
    sah_reset<<<grid, 256, 0, s>>>(ctl, acc);
    sah_cb<<<grid, 256, 0, s>>>(lo[cur], hi[cur], owner[cur], n, acc);
    sah_bin<<<grid, 256, 0, s>>>(lo[cur], hi[cur], owner[cur], n, acc);
    sah_split<<<grid, 128, 0, s>>>(nodes[cur], acc, ctl, nodes[cur ^ 1], small, split, children, parent_inner, parent_leaf, max_depth);
    sah_flag_count<<<tile_grid, kScanBlock, 0, s>>>(lo[cur], hi[cur], owner[cur], split, n, tiles);
    sah_scan_tiles<<<1, 1024, 0, s>>>(tiles, n);
    sah_flag_scan<<<tile_grid, kScanBlock, 0, s>>>(lo[cur], hi[cur], owner[cur], split, n, tiles, S);
    sah_partition<<<grid, 256, 0, s>>>(lo[cur], hi[cur], owner[cur], split, S, n, lo[cur ^ 1], hi[cur ^ 1], owner[cur ^ 1]);
    sah_advance<<<1, 1, 0, s>>>(ctl);
    cur ^= 1;
    ++levels;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if (levels > 64u) return cudaErrorUnknown;
  }
  sah_small<<<grid, kSahSmallWarps * 32, 0, s>>>(small, ctl, lo[cur], hi[cur], children, parent_inner, parent_leaf, max_depth);
  sah_ids<<<grid, 256, 0, s>>>(lo[cur], n, ids);
  // boxes bottom-up (exact unions), leaf refs, depth: the linear BVH's refit over this topology
  lbvh_refit<<<grid, 256, 0, s>>>(d_leaf_box, d_leaf_code, ids, (int)n, children, parent_inner, parent_leaf, arrivals,
                                  static_cast<NodeOut*>(d_inner_out));
  lbvh_depth<<<grid, 256, 0, s>>>((int)n, parent_inner, parent_leaf, deepest);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  unsigned h_deep = 0;
  if ((e = cudaMemcpyAsync(&h_deep, deepest, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  *depth_out = h_deep;
  if (levels_out) *levels_out = levels;
  return cudaSuccess;
}

}  // namespace tutu
