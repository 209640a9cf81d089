// tutu_b200.cu — device side of libtutu_b200.so: context, scene upload, the ray-batch kernels
// and the host loop of the wavefront path tracer, behind the C ABI of include/tutu_b200.h.
// sm_100a only; there is no CPU fallback anywhere in this file.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "tutu_internal.hpp"
#include "wavefront.cuh"
#include "bdpt.cuh"
#include "resident.cuh"
#include "postprocess.cuh"
#include "trace_kernels.hpp"
#include "device_bvh.hpp"
#include <chrono>
#include <cmath>

using namespace tutu;

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
namespace {

struct CudaError {
  cudaError_t code;
  const char* what;
  const char* file;
  int line;
};

#define CUDA_TRY(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) throw CudaError{_e, #expr, __FILE__, __LINE__}; \
  } while (0)

// Every device allocation of the library goes through DevBuf.  A -DTUTU_GUARDS build (libtutu_b200_guard.so,
// tests/test_gpu_guards.py) surrounds each allocation with kGuardBytes of a byte pattern on either side and keeps a
// registry of the live buffers; tutu_debug_guard_check() counts the guard bytes that no longer hold the pattern.
// compute-sanitizer is closed on the GPU pool, so this is the library's own check for writes outside an allocation.
#ifdef TUTU_GUARDS
constexpr size_t kGuardBytes = 64 << 10;
constexpr int kGuardPattern = 0xA5;
std::mutex g_guard_mu;
std::vector<std::pair<void*, size_t>> g_guard_live;  // {first byte after the lower band, bytes}: DevBuf objects may move
uint64_t g_guard_bad_freed = 0;                      // overwritten band bytes of allocations that were freed since
// overwritten bytes in the two bands of one allocation (synchronises the device)
uint64_t guard_bad_bytes(const void* p, size_t bytes) {
  static std::vector<unsigned char> host(kGuardBytes);
  uint64_t bad = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  for (const char* g : {static_cast<const char*>(p) - kGuardBytes, static_cast<const char*>(p) + bytes}) {
    if (cudaMemcpy(host.data(), g, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    for (unsigned char c : host) bad += c != (unsigned char)kGuardPattern;
  }
  return bad;
}
#else
constexpr size_t kGuardBytes = 0;
#endif

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
#ifdef TUTU_GUARDS
      std::lock_guard<std::mutex> lock(g_guard_mu);
      g_guard_bad_freed += guard_bad_bytes(p, bytes);  // a buffer that is replaced by a larger one is checked on its way out
      g_guard_live.erase(std::remove_if(g_guard_live.begin(), g_guard_live.end(), [&](const auto& e) { return e.first == p; }),
                         g_guard_live.end());
#endif
      cudaFree(static_cast<char*>(p) - kGuardBytes);
    }
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    void* base = nullptr;
    CUDA_TRY(cudaMalloc(&base, n + 2 * kGuardBytes));
    p = static_cast<char*>(base) + kGuardBytes;
    bytes = n;
#ifdef TUTU_GUARDS
    // blocking memsets: the guards are in place before any stream can touch the buffer
    CUDA_TRY(cudaMemset(base, kGuardPattern, kGuardBytes));
    CUDA_TRY(cudaMemset(static_cast<char*>(p) + n, kGuardPattern, kGuardBytes));
    std::lock_guard<std::mutex> lock(g_guard_mu);
    g_guard_live.emplace_back(p, n);
#endif
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

template <class T, class A>
void upload_vec(DevBuf& b, const std::vector<T, A>& v, cudaStream_t s) {
  b.ensure(std::max<size_t>(v.size() * sizeof(T), 16));
  if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

}  // namespace

struct WfLane {
  DevBuf pool, ctl, class_perm;
  WfBuffers b{};
  CUtensorMap pool_map{};  // 3-D view {4 floats, entries, arrays} of `pool` for wf_shade's TMA tile fetch
  const void* map_base = nullptr;  // what pool_map was encoded for
  uint64_t map_phys = 0;
  int map_block = 0;
  uint64_t capacity = 0;
  cudaStream_t stream = nullptr;
  WfCtl* ctl_host = nullptr;  // pinned
  cudaEvent_t ev_done = nullptr;
  WfLane() = default;
  WfLane(WfLane&& o) noexcept { *this = std::move(o); }
  WfLane& operator=(WfLane&& o) noexcept {
    std::swap(pool.p, o.pool.p);
    std::swap(pool.bytes, o.pool.bytes);
    std::swap(ctl.p, o.ctl.p);
    std::swap(ctl.bytes, o.ctl.bytes);
    std::swap(class_perm.p, o.class_perm.p);
    std::swap(class_perm.bytes, o.class_perm.bytes);
    b = o.b;
    std::swap(pool_map, o.pool_map);
    std::swap(map_base, o.map_base);
    std::swap(map_phys, o.map_phys);
    std::swap(map_block, o.map_block);
    std::swap(capacity, o.capacity);
    std::swap(stream, o.stream);
    std::swap(ctl_host, o.ctl_host);
    std::swap(ev_done, o.ev_done);
    return *this;
  }
  ~WfLane() {
    if (stream) cudaStreamDestroy(stream);
    if (ctl_host) cudaFreeHost(ctl_host);
    if (ev_done) cudaEventDestroy(ev_done);
  }
};

struct BdptLane {
  DevBuf pool, ctl;
  BdptBuffers b{};
  BdptCtl* ctl_host = nullptr;  // pinned
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_done = nullptr;
  BdptLane() = default;
  BdptLane(const BdptLane&) = delete;
  BdptLane& operator=(const BdptLane&) = delete;
  ~BdptLane() {
    if (stream) cudaStreamDestroy(stream);
    if (ctl_host) cudaFreeHost(ctl_host);
    if (ev_done) cudaEventDestroy(ev_done);
  }
};

constexpr uint64_t kBdptTuneMinBatches = 8;  // shorter renders keep the packet tracer (two batches run alone for the measurement)
constexpr size_t kLocalStackMaxTreeBytes = 32u << 20;  // traversal arrays up to this size: stack in local memory
constexpr int kHostSlotsMax = 4;  // host-buffer ray batches: chunks in flight (trace_host_pipelined)
struct TutuCtx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string last_error;
  std::mutex mu;  // adapters may call UpdateInter from many host threads (PathTracing.hpp:394-429)

  // scene
  bool has_scene = false;
  FlatScene flat;
  DevScene dev{};
  SmallScene small{};  // n = 0 unless the scene has <= kSmallMax primitives
  DevBuf d_inner_fast, d_wide, d_wleaf, d_wbox;
  DevBuf d_leaf_box, d_leaf_code, d_build_scratch;  // inputs and scratch of the device tree build
  int builder_cfg = TUTU_BUILD_AUTO;
  bool device_tree = false;  // d_inner_fast was built on the device (ctx->flat.inner_fast is empty until someone needs it)
  TutuUploadStats upload_stats{};
  WideGrids wide_grids;  // trace_kernels.cu: persistent grids of the wide-tree kernels for the current stack size
  bool use_wide = false;  // tutu_set_traversal_mode(ctx, 6): regular rays walk the compressed 8-wide tree (built on demand)
  DevBuf d_inner, d_geom, d_shade, d_leaftex, d_slot_to_prim, d_materials, d_lights, d_texels;
  DevBuf d_texh[4];
  uint64_t scene_bytes = 0;
  int traversal_mode = 0;

  // ray batches
  DevBuf d_rays, d_hits, d_blocked, d_counts;
  DevBuf d_bin_keys[kHostSlotsMax], d_bin_perm[kHostSlotsMax], d_bin_hist[kHostSlotsMax];  // [pipeline slot]
  // host-buffer ray batches are pipelined in chunks over kHostSlots streams (H2D | walk | D2H overlap)
  cudaStream_t slot_streams[kHostSlotsMax] = {};  // [0] unused (slot 0 runs on `stream`)
  DevBuf d_chunk_rays[kHostSlotsMax], d_chunk_out[kHostSlotsMax];
  // The work cursor and the binning scratch of a slot are shared by every batch traced through it; batches
  // may arrive on different caller streams, so each launch waits for the slot's previous user and marks itself
  // as the last one (the batches of one slot run back to back on the device, whatever their streams).
  cudaEvent_t slot_last_use[kHostSlotsMax] = {};
  int ray_binning = 1;              // 0 = trace in the caller's order
  uint64_t ray_binning_min = 1u << 16;

  // BDPT: two lanes, batches alternate between them
  BdptLane bdpt_lanes[2];

  // Postprocessor scratch: [0..2] intermediates, [3],[4] staging of the host entry point
  DevBuf d_post[5];

  // wavefront
  DevBuf d_accum, d_rgb;
  std::vector<WfLane> wf_lanes;
  uint64_t paths_in_flight_cfg = 0;  // per lane
  int lanes_cfg = 0;
  int grid_lanes = 0;
  bool sort_by_class = false;  // wavefront.cuh: wf_classify (scenes with more than one shading class)
  int shade_block = TUTU_SHADE_BLOCK;  // wavefront.cuh: kShadeBlockSimple for all-Lambertian untextured scenes
  int grid_shade_block = 0;
  size_t grid_stack_smem = 0;
  bool grid_small = false;
  bool grid_wide = false;
  bool grid_stack_shared = false;
  // Traversal stack of the tree kernels: shared memory for trees that do not fit near the SM, local memory for small
  // ones (decided at upload from the bytes of the traversal arrays)
  bool stack_shared = true;
  // BDPT queue tracers (q_extend / q_shadow_add on scenes with a tree): one packet of rays at a time, or persistent lanes with
  // phase-separated steps (trace.cuh: trace_queue_lanes).  Which one is faster depends on the scene, so the first
  // render of at least kBdptTuneMinBatches batches times one batch with each and keeps the winner for the scene.
  int bdpt_tracer_cfg = 0;     // tutu_bdpt_queue_tracer: 0 = measured per scene, 1 = packets, 2 = persistent lanes
  int bdpt_tracer_tuned = -1;  // -1 = not measured yet for the uploaded scene, else 0 = packets / 1 = lanes
  float bdpt_tune_ms[2] = {0.f, 0.f};
  int stack_cfg = 0;  // tutu_traversal_stack: 0 = by tree size, 1 = shared memory, 2 = local memory
  size_t tree_bytes = 0;  // traversal arrays of the uploaded scene (both topologies' nodes + leaf geometry)
  int profile_stages = 0;
  // 0 = automatic (register-resident kernel for small renders of scenes that fit the constant bank, else
  // wavefront), 1 = wavefront, 2 = register-resident (fails on scenes that do not fit)
  int pipeline_cfg = 0;
  DevBuf d_resident_ctl;
  int grid_resident = 0;
  TutuRenderStats stats{};
  int grid_extend = 0, grid_shade = 0, grid_shadow = 0, grid_raygen = 0;
};

namespace {

const char* ctx_error_hook(const TutuCtx* c) { return c->last_error.c_str(); }
struct HookInstaller {
  HookInstaller() { g_ctx_error_hook = &ctx_error_hook; }
} g_hook_installer;

int fail(TutuCtx* ctx, int code, const std::string& msg) {
  set_error(msg);
  if (ctx) ctx->last_error = msg;
  return code;
}
int fail_cuda(TutuCtx* ctx, const CudaError& e) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e.code, cudaGetErrorString(e.code),
           e.file, e.line, e.what);
  cudaGetLastError();  // clear the sticky-less error state
  return fail(ctx, TUTU_E_CUDA, buf);
}

#define API_BEGIN(ctx)                                                 \
  if (!(ctx)) return fail(nullptr, TUTU_E_INVALID, "null context");    \
  std::lock_guard<std::mutex> _lock((ctx)->mu);                        \
  try {                                                                \
    CUDA_TRY(cudaSetDevice((ctx)->device));
#define API_END(ctx)                                                   \
  }                                                                    \
  catch (const CudaError& e) { return fail_cuda((ctx), e); }           \
  catch (const std::bad_alloc&) { return fail((ctx), TUTU_E_NOMEM, "out of host memory"); } \
  catch (const std::exception& e) { return fail((ctx), TUTU_E_INVALID, std::string("unexpected exception: ") + e.what()); } \
  catch (...) { return fail((ctx), TUTU_E_INVALID, "unexpected exception"); }

// ---------------------------------------------------------------------------------------------
// ray-batch kernels
// ---------------------------------------------------------------------------------------------
// Warps pull 32-ray packets from a global cursor (persistent threads: grid = SMs x resident blocks).
// MODE 0: ordered + pruned walk, FMNMX slab test for regular rays (production).
// MODE 1: the literal reference walk (tests only).
// MODE 2: per-lane ray refill between traversal rounds (trace_persistent; measured slower than
//         MODE 0 on every workload, kept for the record — DESIGN.md §5).
// ---- ray binning ----------------------------------------------------------------------------
// A caller's ray batch arrives in arbitrary order; a warp of 32 unrelated rays walks 32 unrelated
// root-to-leaf paths (ncu on the 2^24-ray batch: 10 of 32 lanes active per instruction, L1/TEX at
// 86 % of peak from divergent 64-byte node fetches).  Large batches are therefore traced in a
// coherent ORDER: rays are counting-sorted by a key made of the Morton code of the cell where the
// ray enters the scene box and of an octahedral direction bin, and the tracer walks a permutation.
// Each ray's result is computed exactly as before and written to its own slot, so the output is
// independent of the order (tests compare sorted vs unsorted bit for bit).
constexpr int kBinOriginBits = 5;  // per axis, for the largest batches; small batches use fewer (bin_origin_bits)
constexpr int kBinDirBits = 3;     // per octahedral axis
constexpr int kBinKeyBits = 3 * kBinOriginBits + 2 * kBinDirBits;
constexpr unsigned kBinCount = 1u << kBinKeyBits;  // capacity of the histogram buffer
// Origin bits per axis for a batch of n rays: about 8 rays per bin (2^24 rays -> 5 bits = 2^21 bins; 2^16 rays -> 2
// bits = 2^12 bins), so that clearing and scanning the histogram stays small next to the walk for small batches.
inline int bin_origin_bits(uint64_t n) {
  int lg = 0;
  while ((1ull << (lg + 1)) <= n) ++lg;
  return std::min(kBinOriginBits, std::max(2, (lg - 9) / 3));
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__device__ __forceinline__ unsigned ray_bin_key(const DevScene& sc, const float4 o, const float4 d, int origin_bits) {
  // entry point into the root box (plain fp32: ordering only, never a hit decision)
  const float ix = 1.f / d.x, iy = 1.f / d.y, iz = 1.f / d.z;
  const float ax = (sc.root_lo[0] - o.x) * ix, bx = (sc.root_hi[0] - o.x) * ix;
  const float ay = (sc.root_lo[1] - o.y) * iy, by = (sc.root_hi[1] - o.y) * iy;
  const float az = (sc.root_lo[2] - o.z) * iz, bz = (sc.root_hi[2] - o.z) * iz;
  float te = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
  if (!(te < 3.0e38f)) te = 0.f;
  const float px = o.x + te * d.x, py = o.y + te * d.y, pz = o.z + te * d.z;
  const float cells = (float)(1 << origin_bits);
  auto q = [&](float p, float lo, float hi) {
    const float w = hi - lo;
    float u = w > 0.f ? (p - lo) / w : 0.f;
    u = fminf(fmaxf(u, 0.f), 0.999999f);
    return (unsigned)(u * cells);
  };
  const unsigned cell = spread3(q(px, sc.root_lo[0], sc.root_hi[0])) | (spread3(q(py, sc.root_lo[1], sc.root_hi[1])) << 1) |
                        (spread3(q(pz, sc.root_lo[2], sc.root_hi[2])) << 2);
  // octahedral direction bin
  const float l1 = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
  float u = l1 > 0.f ? d.x / l1 : 0.f, v = l1 > 0.f ? d.z / l1 : 0.f;
  if (d.y < 0.f) {
    const float uu = (1.f - fabsf(v)) * (u >= 0.f ? 1.f : -1.f), vv = (1.f - fabsf(u)) * (v >= 0.f ? 1.f : -1.f);
    u = uu, v = vv;
  }
  const float dcells = (float)(1 << kBinDirBits);
  const unsigned du = (unsigned)(fminf(fmaxf(u * 0.5f + 0.5f, 0.f), 0.999999f) * dcells);
  const unsigned dv = (unsigned)(fminf(fmaxf(v * 0.5f + 0.5f, 0.f), 0.999999f) * dcells);
  const unsigned dirbin = (du << kBinDirBits) | dv;
  return (dirbin << (3 * origin_bits)) | (cell & ((1u << (3 * origin_bits)) - 1u));
}

__global__ void __launch_bounds__(256)
k_bin_count(const __grid_constant__ DevScene sc, const float4* __restrict__ rays, unsigned n, int origin_bits,
            unsigned* __restrict__ keys, unsigned* __restrict__ hist) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned k = ray_bin_key(sc, __ldg(rays + 2 * (size_t)i), __ldg(rays + 2 * (size_t)i + 1), origin_bits);
    keys[i] = k;
    atomicAdd(hist + k, 1u);
  }
}

// exclusive scan of the histogram in three coalesced passes: per-chunk totals, scan of the totals,
// per-chunk exclusive scan + chunk offset.  A chunk = 2048 bins = 256 threads x 8 bins.
constexpr unsigned kScanChunk = 2048u;
constexpr unsigned kScanChunks = kBinCount / kScanChunk;
static_assert(kScanChunks <= 1024u && kBinCount % kScanChunk == 0u, "scan layout");

__device__ __forceinline__ unsigned block_exclusive_scan_256(unsigned v, unsigned* total) {
  __shared__ unsigned ws[8];
  const unsigned t = threadIdx.x, lane = t & 31u;
  unsigned incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (unsigned)o) incl += y;
  }
  if (lane == 31u) ws[t >> 5] = incl;
  __syncthreads();
  unsigned before = 0, all = 0;
  for (unsigned w = 0; w < 8; ++w) {
    if (w < (t >> 5)) before += ws[w];
    all += ws[w];
  }
  __syncthreads();
  *total = all;
  return before + incl - v;
}

__global__ void __launch_bounds__(256)
k_bin_scan_totals(const unsigned* __restrict__ hist, unsigned* __restrict__ totals) {
  const uint4* h = reinterpret_cast<const uint4*>(hist + (size_t)blockIdx.x * kScanChunk) + 2 * threadIdx.x;
  const uint4 a = h[0], b = h[1];
  unsigned total;
  block_exclusive_scan_256(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w, &total);
  if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
k_bin_scan_chunks(unsigned* __restrict__ totals, unsigned n_chunks) {
  __shared__ unsigned ws[32];
  const unsigned t = threadIdx.x, lane = t & 31u;
  const unsigned v = t < n_chunks ? totals[t] : 0u;
  unsigned incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= (unsigned)o) incl += y;
  }
  if (lane == 31u) ws[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    unsigned w = ws[t];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xFFFFFFFFu, w, o);
      if (t >= (unsigned)o) w += y;
    }
    ws[t] = w;
  }
  __syncthreads();
  if (t < n_chunks) totals[t] = incl - v + ((t >> 5) ? ws[(t >> 5) - 1] : 0u);
}

__global__ void __launch_bounds__(256)
k_bin_scan_apply(unsigned* __restrict__ hist, const unsigned* __restrict__ totals) {
  uint4* h = reinterpret_cast<uint4*>(hist + (size_t)blockIdx.x * kScanChunk) + 2 * threadIdx.x;
  const uint4 a = h[0], b = h[1];
  unsigned total;
  unsigned base = block_exclusive_scan_256(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w, &total) + totals[blockIdx.x];
  uint4 oa, ob;
  oa.x = base, base += a.x;
  oa.y = base, base += a.y;
  oa.z = base, base += a.z;
  oa.w = base, base += a.w;
  ob.x = base, base += b.x;
  ob.y = base, base += b.y;
  ob.z = base, base += b.z;
  ob.w = base;
  h[0] = oa, h[1] = ob;
}

__global__ void __launch_bounds__(256)
k_bin_scatter(const unsigned* __restrict__ keys, unsigned n, unsigned* __restrict__ offsets, unsigned* __restrict__ perm) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    perm[atomicAdd(offsets + keys[i], 1u)] = i;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_trace_closest(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss,
                const float4* __restrict__ rays, unsigned long long n, TutuHit* __restrict__ out,
                unsigned long long* __restrict__ next, const unsigned* __restrict__ perm = nullptr) {
  extern __shared__ unsigned long long s_stack[];  // MODE 0: traversal stack (trace.cuh: SharedStack)
  auto store = [&](unsigned long long i, const Hit& h) {
    const int prim = h.slot >= 0 ? __ldg(sc.slot_to_prim + (h.slot & (int)kSlotMask)) : -1;
    reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), h.t, h.u, h.v);
  };
#ifdef TUTU_EXPERIMENTS
  if (MODE == 2) {
    trace_persistent<false>(
        sc, n, next,
        [&](unsigned long long i, Ray& r, float& dis) {
          const float4 o = __ldg(rays + 2 * i);
          const float4 d = __ldg(rays + 2 * i + 1);
          r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
          dis = 0.f;
        },
        [&](unsigned long long i, const Walk& w) { store(i, w.best); });
    return;
  }
#endif
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(next, 32ull);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
    const unsigned long long j = base + lane;
    if (MODE == 0) {  // production: the warp walks together, primitive tests batched (trace.cuh: traverse_batched)
      const bool valid = j < n;
      const unsigned long long i = valid ? (perm ? (unsigned long long)__ldg(perm + j) : j) : 0ull;
      float4 o = make_float4(0, 0, 0, 0), d = make_float4(1, 0, 0, 0);
      if (valid) o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
      Hit h;
      traverse_batched<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, valid, h, s_stack);
      if (valid) store(i, h);
      continue;
    }
    if (j < n) {
      const unsigned long long i = perm ? (unsigned long long)__ldg(perm + j) : j;
      const float4 o = __ldg(rays + 2 * i);
      const float4 d = __ldg(rays + 2 * i + 1);
      Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
      Hit h;
      if (MODE == 3)
        traverse_small<false>(sc, ss, r, 0.f, h);
      else if (MODE == 0)
        traverse_shared<false>(sc, r, 0.f, h, s_stack);
      else
        traverse<false, 1, false>(sc, r, 0.f, h, nullptr);
      store(i, h);
    }
    __syncwarp();
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_trace_any(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss,
            const float4* __restrict__ rays, unsigned long long n, uint8_t* __restrict__ out,
            unsigned long long* __restrict__ next, const unsigned* __restrict__ perm = nullptr) {
  extern __shared__ unsigned long long s_stack[];  // MODE 0: traversal stack, 32-bit entries (trace.cuh: SharedStack<true>)
#ifdef TUTU_EXPERIMENTS
  if (MODE == 2) {
    trace_persistent<true>(
        sc, n, next,
        [&](unsigned long long i, Ray& r, float& dis) {
          const float4 o = __ldg(rays + 2 * i);
          const float4 d = __ldg(rays + 2 * i + 1);
          r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
          dis = d.w;
        },
        [&](unsigned long long i, const Walk& w) { out[i] = w.best.slot >= 0 ? 1 : 0; });
    return;
  }
#endif
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(next, 32ull);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
    const unsigned long long j = base + lane;
    if (j < n) {
      const unsigned long long i = perm ? (unsigned long long)__ldg(perm + j) : j;
      const float4 o = __ldg(rays + 2 * i);
      const float4 d = __ldg(rays + 2 * i + 1);
      Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
      Hit h;
      if (MODE == 3)
        out[i] = traverse_small<true>(sc, ss, r, d.w, h) ? 1 : 0;
      else if (MODE == 0)
        out[i] = traverse_shared<true>(sc, r, d.w, h, s_stack) ? 1 : 0;
      else if (MODE == 4)
        out[i] = traverse_structured<true>(sc, r, d.w, h) ? 1 : 0;
      else
        out[i] = traverse<true, 1, false>(sc, r, d.w, h, nullptr) ? 1 : 0;
    }
    __syncwarp();
  }
}

#ifdef TUTU_EXPERIMENTS
// experimental walk flavours behind tutu_set_traversal_mode(ctx, 10 + VARIANT): 32-ray packets
template <bool ANY, int VARIANT>
__global__ void __launch_bounds__(256)
k_trace_variant(const __grid_constant__ DevScene sc, const float4* __restrict__ rays, unsigned long long n, TutuHit* __restrict__ out,
                uint8_t* __restrict__ out_any, unsigned long long* __restrict__ next, const unsigned* __restrict__ perm) {
  if (VARIANT == 6) {  // persistent lanes with refill, shared-memory stack
    extern __shared__ unsigned long long s_stack[];
    trace_refill<ANY>(
        sc, n, next, s_stack,
        [&](unsigned long long j, Ray& r, float& dis) {
          const unsigned long long i = perm ? (unsigned long long)__ldg(perm + j) : j;
          const float4 o = __ldg(rays + 2 * i);
          const float4 d = __ldg(rays + 2 * i + 1);
          r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
          dis = ANY ? d.w : 0.f;
        },
        [&](unsigned long long j, const Walk& w) {
          const unsigned long long i = perm ? (unsigned long long)__ldg(perm + j) : j;
          if (ANY) {
            out_any[i] = w.best.slot >= 0 ? 1 : 0;
          } else {
            const int prim = w.best.slot >= 0 ? __ldg(sc.slot_to_prim + (w.best.slot & (int)kSlotMask)) : -1;
            reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), w.best.t, w.best.u, w.best.v);
          }
        });
    return;
  }
  const unsigned lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(next, 32ull);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
    const unsigned long long j = base + lane;
    const bool valid = j < n;
    const unsigned long long i = valid ? (perm ? (unsigned long long)__ldg(perm + j) : j) : 0ull;
    if (VARIANT == 7) {  // warp walk with batched leaf tests
      extern __shared__ unsigned long long s_stack[];
      float4 o = make_float4(0, 0, 0, 0), d = make_float4(1, 0, 0, 0);
      if (valid) o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
      Hit h;
      const bool any = traverse_batched<ANY>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, ANY ? d.w : 0.f, valid, h, s_stack);
      if (valid) {
        if (ANY) {
          out_any[i] = any ? 1 : 0;
        } else {
          const int prim = h.slot >= 0 ? __ldg(sc.slot_to_prim + (h.slot & (int)kSlotMask)) : -1;
          reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), h.t, h.u, h.v);
        }
      }
      continue;
    }
    if (VARIANT == 5) {
      extern __shared__ unsigned long long s_stack[];
      if (valid) {
        const float4 o = __ldg(rays + 2 * i);
        const float4 d = __ldg(rays + 2 * i + 1);
        Hit h;
        const bool any = traverse_shared<ANY>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, ANY ? d.w : 0.f, h, s_stack);
        if (ANY) {
          out_any[i] = any ? 1 : 0;
        } else {
          const int prim = h.slot >= 0 ? __ldg(sc.slot_to_prim + (h.slot & (int)kSlotMask)) : -1;
          reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), h.t, h.u, h.v);
        }
      }
      __syncwarp();
      continue;
    }
    if (VARIANT == 4) {
      float4 o = make_float4(0, 0, 0, 0), d = make_float4(1, 0, 0, 0);
      if (valid) o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
      Hit h;
      const bool any = traverse_warp<ANY>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, ANY ? d.w : 0.f, valid, h);
      if (valid) {
        if (ANY) {
          out_any[i] = any ? 1 : 0;
        } else {
          const int prim = h.slot >= 0 ? __ldg(sc.slot_to_prim + (h.slot & (int)kSlotMask)) : -1;
          reinterpret_cast<float4*>(out)[i] = make_float4(__int_as_float(prim), h.t, h.u, h.v);
        }
      }
      continue;
    }
    if (valid) {
      const float4 o = __ldg(rays + 2 * i);
      const float4 d = __ldg(rays + 2 * i + 1);
      Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
      Hit h;
      const bool any = traverse_variant<ANY, (VARIANT >= 4 ? 1 : VARIANT)>(sc, r, d.w, h);  // (4..7 return above)
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

#endif  // TUTU_EXPERIMENTS

template <bool ANY>
__global__ void __launch_bounds__(256)
k_trace_count(const __grid_constant__ DevScene sc, const float4* __restrict__ rays, unsigned long long n,
              unsigned long long* __restrict__ counts) {
  unsigned long long nodes = 0, prims = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float4 o = __ldg(rays + 2 * i);
    const float4 d = __ldg(rays + 2 * i + 1);
    Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
    Hit h;
    VisitCount vc;
    traverse<ANY, 0, true>(sc, r, d.w, h, &vc);
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

template <class K>
int persistent_grid(TutuCtx* ctx, K kernel, int block, size_t dyn_smem = 0) {
  int per_sm = 0;
  if (dyn_smem) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem));
  if (per_sm < 1) per_sm = 1;
#ifdef TUTU_EXPERIMENTS
  if (const char* e = getenv("TUTU_GRID_DIV")) per_sm = std::max(1, per_sm / std::max(1, atoi(e)));  // room for a co-running kernel
#endif
  return ctx->sm_count * per_sm;  // a multiple of the SM count: one resident wave
}

// shared-memory stack bytes for a block: one 64-bit word per thread and level; a ray pushes at
// most one entry per tree level, +1 slack
size_t stack_smem(const TutuCtx* ctx, int block, bool any) {
  return (size_t)block * (std::max(ctx->flat.depth, ctx->flat.depth_fast) + 1) * (any ? sizeof(unsigned) : sizeof(unsigned long long));
}

size_t wide_stack_smem(const TutuCtx* ctx, int block) { return (size_t)block * (ctx->flat.wide_depth + 1) * sizeof(unsigned long long); }
// grids of the wide-tree kernels (trace_kernels.cu) for the uploaded scene's stack size
const WideGrids& wide_grids_for(TutuCtx* ctx) {
  const size_t sm = wide_stack_smem(ctx, kTraceBlock);
  if (ctx->wide_grids.smem != sm) CUDA_TRY(wide_grids(ctx->sm_count, sm, &ctx->wide_grids));
  return ctx->wide_grids;
}

// Orders the stream after the previous batch that used the slot's cursor / binning scratch.
void slot_acquire(TutuCtx* ctx, int slot, cudaStream_t s) {
  if (!ctx->slot_last_use[slot])
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->slot_last_use[slot], cudaEventDisableTiming));
  else
    CUDA_TRY(cudaStreamWaitEvent(s, ctx->slot_last_use[slot], 0));
}
void slot_release(TutuCtx* ctx, int slot, cudaStream_t s) { CUDA_TRY(cudaEventRecord(ctx->slot_last_use[slot], s)); }

// Builds the coherent traversal order of a batch (nullptr = trace in the caller's order).
const unsigned* bin_rays(TutuCtx* ctx, const float4* rays, uint64_t n, cudaStream_t s, int slot = 0) {
  if (!ctx->ray_binning || !(ctx->traversal_mode == 0 || ctx->traversal_mode >= 10) || ctx->small.n > 0 || n < ctx->ray_binning_min || n >= (1ull << 32))
    return nullptr;
  ctx->d_bin_keys[slot].ensure(n * sizeof(unsigned));
  ctx->d_bin_perm[slot].ensure(n * sizeof(unsigned));
  ctx->d_bin_hist[slot].ensure((size_t)(kBinCount + 1024) * sizeof(unsigned));
  unsigned* keys = ctx->d_bin_keys[slot].as<unsigned>();
  unsigned* perm = ctx->d_bin_perm[slot].as<unsigned>();
  unsigned* hist = ctx->d_bin_hist[slot].as<unsigned>();
  const int origin_bits = bin_origin_bits(n);
  const unsigned bins = 1u << (3 * origin_bits + 2 * kBinDirBits), chunks = bins / kScanChunk;  // only the bins in use
  CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)bins * sizeof(unsigned), s));
  const int grid = ctx->sm_count * 8;
  k_bin_count<<<grid, 256, 0, s>>>(ctx->dev, rays, (unsigned)n, origin_bits, keys, hist);
  k_bin_scan_totals<<<chunks, 256, 0, s>>>(hist, hist + kBinCount);
  k_bin_scan_chunks<<<1, 1024, 0, s>>>(hist + kBinCount, chunks);
  k_bin_scan_apply<<<chunks, 256, 0, s>>>(hist, hist + kBinCount);
  k_bin_scatter<<<grid, 256, 0, s>>>(keys, (unsigned)n, hist, perm);
  CUDA_TRY(cudaGetLastError());
  return perm;
}

void launch_closest(TutuCtx* ctx, const float* d_rays, uint64_t n, TutuHit* d_out, cudaStream_t s, int slot = 0) {
  if (n == 0) return;
  ctx->d_counts.ensure(256);
  unsigned long long* next = ctx->d_counts.as<unsigned long long>() + 4 + 2 * slot;  // cursors: [4] closest, [5] any, [6],[7] slot 1
  slot_acquire(ctx, slot, s);
  CUDA_TRY(cudaMemsetAsync(next, 0, sizeof(unsigned long long), s));
  const float4* rays = reinterpret_cast<const float4*>(d_rays);
#ifdef TUTU_EXPERIMENTS
  if (ctx->traversal_mode >= 10) {
#define TUTU_VAR_C(V)                                                                       \
  case V: {                                                                                 \
    const size_t sm = V >= 5 ? stack_smem(ctx, 256, false) : 0;                             \
    int grid = persistent_grid(ctx, k_trace_variant<false, V>, 256, sm);                    \
    k_trace_variant<false, V><<<grid, 256, sm, s>>>(ctx->dev, rays, n, d_out, nullptr, next, perm); \
  } break;
    const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
    switch (ctx->traversal_mode - 10) {
      TUTU_VAR_C(0) TUTU_VAR_C(1) TUTU_VAR_C(2) TUTU_VAR_C(3) TUTU_VAR_C(4) TUTU_VAR_C(5) TUTU_VAR_C(6) TUTU_VAR_C(7)
    }
#undef TUTU_VAR_C
  } else if (ctx->traversal_mode == 2) {
    int grid = persistent_grid(ctx, k_trace_closest<2>, 256);
    k_trace_closest<2><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
  } else
#endif
  if (ctx->traversal_mode == 1) {
    int grid = persistent_grid(ctx, k_trace_closest<1>, 256);
    k_trace_closest<1><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
  } else {
    if (ctx->small.n > 0 && ctx->traversal_mode == 0) {
      int grid = persistent_grid(ctx, k_trace_closest<3>, 256);
      k_trace_closest<3><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
    } else if (ctx->dev.wide) {
      const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
      const WideGrids& g = wide_grids_for(ctx);
      CUDA_TRY(wide_launch_batch(false, g.batch_closest, g.smem, s, ctx->dev, rays, n, d_out, nullptr, next, perm));
    } else {
      const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
      const size_t sm = stack_smem(ctx, 256, false);
      int grid = persistent_grid(ctx, k_trace_closest<0>, 256, sm);
      k_trace_closest<0><<<grid, 256, sm, s>>>(ctx->dev, ctx->small, rays, n, d_out, next, perm);
    }
  }
  CUDA_TRY(cudaGetLastError());
  slot_release(ctx, slot, s);
}

void launch_any(TutuCtx* ctx, const float* d_rays, uint64_t n, uint8_t* d_out, cudaStream_t s, int slot = 0) {
  if (n == 0) return;
  ctx->d_counts.ensure(256);
  unsigned long long* next = ctx->d_counts.as<unsigned long long>() + 5 + 2 * slot;
  slot_acquire(ctx, slot, s);
  CUDA_TRY(cudaMemsetAsync(next, 0, sizeof(unsigned long long), s));
  const float4* rays = reinterpret_cast<const float4*>(d_rays);
#ifdef TUTU_EXPERIMENTS
  if (ctx->traversal_mode >= 10) {
#define TUTU_VAR_A(V)                                                                      \
  case V: {                                                                                \
    const size_t sm = V >= 5 ? stack_smem(ctx, 256, true) : 0;                             \
    int grid = persistent_grid(ctx, k_trace_variant<true, V>, 256, sm);                    \
    k_trace_variant<true, V><<<grid, 256, sm, s>>>(ctx->dev, rays, n, nullptr, d_out, next, perm); \
  } break;
    const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
    switch (ctx->traversal_mode - 10) {
      TUTU_VAR_A(0) TUTU_VAR_A(1) TUTU_VAR_A(2) TUTU_VAR_A(3) TUTU_VAR_A(4) TUTU_VAR_A(5) TUTU_VAR_A(6) TUTU_VAR_A(7)
    }
#undef TUTU_VAR_A
  } else if (ctx->traversal_mode == 2) {
    int grid = persistent_grid(ctx, k_trace_any<2>, 256);
    k_trace_any<2><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
  } else
#endif
  if (ctx->traversal_mode == 1) {
    int grid = persistent_grid(ctx, k_trace_any<1>, 256);
    k_trace_any<1><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
  } else {
    if (ctx->small.n > 0 && ctx->traversal_mode == 0) {
      int grid = persistent_grid(ctx, k_trace_any<3>, 256);
      k_trace_any<3><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next);
    } else if (ctx->dev.wide) {
      const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
      const WideGrids& g = wide_grids_for(ctx);
      CUDA_TRY(wide_launch_batch(true, g.batch_any, g.smem, s, ctx->dev, rays, n, nullptr, d_out, next, perm));
    } else {
      const unsigned* perm = bin_rays(ctx, rays, n, s, slot);
      if (ctx->stack_shared) {
        const size_t sm = stack_smem(ctx, 256, true);
        int grid = persistent_grid(ctx, k_trace_any<0>, 256, sm);
        k_trace_any<0><<<grid, 256, sm, s>>>(ctx->dev, ctx->small, rays, n, d_out, next, perm);
      } else {
        int grid = persistent_grid(ctx, k_trace_any<4>, 256);
        k_trace_any<4><<<grid, 256, 0, s>>>(ctx->dev, ctx->small, rays, n, d_out, next, perm);
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  slot_release(ctx, slot, s);
}

// staging buffers of wf_shade (wavefront.cuh); scenes shaded through class lists gather their records and need none
size_t shade_stage_smem(const TutuCtx* ctx) {
  return ctx->sort_by_class ? 0 : (size_t)ctx->shade_block * kShadeStages * 7 * sizeof(float4);
}
// small-scene kernels (flat leaf-box tests) unless traversal mode 4 forces the general tree walk
bool use_small(const TutuCtx* ctx) { return ctx->small.n > 0 && ctx->traversal_mode != 4; }

int check_scene(TutuCtx* ctx) {
  if (!ctx->has_scene) return fail(ctx, TUTU_E_STATE, "no scene uploaded (call tutu_scene_upload first)");
  return TUTU_OK;
}
int check_pipeline(TutuCtx* ctx) {
  if (ctx->pipeline_cfg == 2 && !use_small(ctx))
    return fail(ctx, TUTU_E_STATE, "the register-resident pipeline needs a scene of at most 32 primitives");
  return TUTU_OK;
}

void fill_raygen(const FlatScene& f, RayGenK* k) {
  memcpy(k->eye, f.raygen.eye, 12);
  memcpy(k->ul, f.raygen.ul, 12);
  memcpy(k->dh, f.raygen.delta_h, 12);
  memcpy(k->dv, f.raygen.delta_v, 12);
  memcpy(k->coh, f.raygen.c_off_h, 12);
  memcpy(k->cov, f.raygen.c_off_v, 12);
  k->width = f.raygen.width;
  k->height = f.raygen.height;
}

// ---------------------------------------------------------------------------------------------
// wavefront host loop
// ---------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
void encode_pool_map(CUtensorMap* map, void* pool, uint64_t entries, uint64_t arrays, int block) {
  using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) throw CudaError{cudaErrorNotSupported, "cuTensorMapEncodeTiled (driver entry point)", __FILE__, __LINE__};
    encode = reinterpret_cast<Encode>(fn);
  }
  // {floats of a run of `run` entries (<= 256, the box limit), runs, arrays}: the box rows are 1 KB of contiguous memory
  const cuuint64_t run = std::min(block, 64);
  const cuuint64_t dims[3] = {4 * run, entries / run, arrays};
  const cuuint64_t strides[2] = {run * sizeof(float4), entries * sizeof(float4)};  // bytes between runs, between arrays
  const cuuint32_t box[3] = {(cuuint32_t)(4 * run), (cuuint32_t)(block / run), 7};
  const cuuint32_t elem[3] = {1, 1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, pool, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError{cudaErrorInvalidValue, "cuTensorMapEncodeTiled (queue pool view)", __FILE__, __LINE__};
}

// One wavefront "lane": its own queues, control block and stream.  Lanes run interleaved on separate
// streams, staggered by one stage, every kernel at its full persistent grid; the
// default is two lanes of 16 Mi paths.
void lane_prepare(TutuCtx* ctx, WfLane& L, uint64_t cap) {
  cap = (cap + 255) & ~(uint64_t)255;
  if (cap > L.capacity) L.capacity = cap;
  cap = L.capacity;
  // wf_shade's blocks hold up to kAppendIters x blockDim reserved entries each (wavefront.cuh: queue appends), so a queue of
  // `cap` live entries can extend this far past `cap`; WfBuffers::capacity stays the number of paths in flight
  const uint64_t shading_blocks = std::min<uint64_t>((uint64_t)std::max(ctx->grid_shade, 1), cap / ctx->shade_block + 1);
  const uint64_t slack = (shading_blocks * kAppendIters * ctx->shade_block + 255) & ~(uint64_t)255;
  const uint64_t phys = cap + slack;
  const size_t n_arrays = 2 * 6 + 1 + 4;  // 2*(6 queues) + hit + 4 shadow arrays, float4 each
  L.pool.ensure(n_arrays * phys * sizeof(float4));
  float4* p = L.pool.as<float4>();
  WfBuffers& b = L.b;
  for (int k = 0; k < 2; ++k) {  // set 0 | hit | set 1: the seven arrays wf_shade reads are adjacent for either k
    b.ray_o[k] = p, p += phys;
    b.ray_d[k] = p, p += phys;
    b.st0[k] = p, p += phys;
    b.st1[k] = p, p += phys;
    b.st2[k] = p, p += phys;
    b.st3[k] = p, p += phys;
    if (k == 0) b.hit = p, p += phys;
  }
  if (L.map_base != L.pool.p || L.map_phys != phys || L.map_block != ctx->shade_block) {
    encode_pool_map(&L.pool_map, L.pool.p, phys, n_arrays, ctx->shade_block);
    L.map_base = L.pool.p, L.map_phys = phys, L.map_block = ctx->shade_block;
  }
  b.sh_o = p, p += phys;
  b.sh_d = p, p += phys;
  b.sh_c = p, p += phys;
  b.sh_L = p, p += phys;
  L.ctl.ensure(sizeof(WfCtl));
  b.ctl = L.ctl.as<WfCtl>();
  b.capacity = (unsigned)cap;
  b.class_perm = nullptr;
  if (ctx->sort_by_class) {
    L.class_perm.ensure((size_t)kShadeClasses * cap * sizeof(unsigned));
    b.class_perm = L.class_perm.as<unsigned>();
  }
  if (!L.stream) CUDA_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
  if (!L.ctl_host) CUDA_TRY(cudaMallocHost(&L.ctl_host, sizeof(WfCtl)));
  if (!L.ev_done) CUDA_TRY(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
}

struct StageTimer {
  bool on;
  std::vector<cudaEvent_t> ev;
  std::vector<int> tag;  // stage id of the interval that STARTS at event k
  explicit StageTimer(bool enable) : on(enable) {}
  StageTimer(StageTimer&&) = default;
  StageTimer(const StageTimer&) = delete;
  ~StageTimer() {
    for (auto e : ev) cudaEventDestroy(e);
  }
  void mark(int stage, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(e, s));
    ev.push_back(e);
    tag.push_back(stage);
  }
  void resolve(TutuRenderStats* st) {
    if (!on) return;
    for (size_t k = 0; k + 1 < ev.size(); ++k) {
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      switch (tag[k]) {
        case 1: st->extend_ms += ms; break;
        case 2: st->shade_ms += ms; break;
        case 3: st->shadow_ms += ms; break;
        default: st->other_ms += ms; break;
      }
    }
  }
};

// resident.cuh: the whole render is one persistent launch on `s` (plus the 40-byte control block).
void resident_render(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count, uint64_t seed, float* d_accum,
                     cudaStream_t s) {
  const FlatScene& f = ctx->flat;
  const uint64_t npix = (uint64_t)f.raygen.width * f.raygen.height;
  ctx->stats = TutuRenderStats{};
  if (npix * sample_count == 0) return;
  if (!ctx->grid_resident) {
    CUDA_TRY(pt_resident_grid(ctx->sm_count, &ctx->grid_resident));
  }
  const uint64_t threads = (uint64_t)ctx->grid_resident * TUTU_RESIDENT_BLOCK;
  // samples per work item: 64, less when the frame is too small to give every lane ~4 items
  uint32_t chunk = 64;
#ifdef TUTU_EXPERIMENTS
  if (const char* e = getenv("TUTU_RESIDENT_CHUNK")) chunk = (uint32_t)std::max(1, atoi(e));
#endif
  while (chunk > 1 && npix * ((sample_count + chunk - 1) / chunk) < 4 * threads) chunk /= 2;
  chunk = std::min(chunk, sample_count);
  ResidentArgs a{};
  fill_raygen(f, &a.rk);
  a.sample_begin = sample_begin;
  a.sample_count = sample_count;
  a.chunk = chunk;
  a.n_items = npix * ((sample_count + chunk - 1) / chunk);
  a.seed = seed;
  a.accum = d_accum;
  ctx->d_resident_ctl.ensure(sizeof(ResidentCtl));
  a.ctl = ctx->d_resident_ctl.as<ResidentCtl>();
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  try {
    CUDA_TRY(cudaEventRecord(e0, s));
    CUDA_TRY(cudaMemsetAsync(a.ctl, 0, sizeof(ResidentCtl), s));
    const uint64_t want = (a.n_items + TUTU_RESIDENT_BLOCK - 1) / TUTU_RESIDENT_BLOCK;
    const int grid = (int)std::min<uint64_t>((uint64_t)ctx->grid_resident, std::max<uint64_t>(want, 1));
    CUDA_TRY(pt_resident_launch(grid, s, ctx->dev, ctx->small, a));
    ResidentCtl h{};
    CUDA_TRY(cudaMemcpyAsync(&h, a.ctl, sizeof(h), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    TutuRenderStats& st = ctx->stats;
    st.gpu_ms = ms;
    st.paths = npix * sample_count;
    st.extend_rays = h.sum_extend;
    st.shadow_rays = h.sum_shadow;
    st.shade_calls = h.sum_extend;
    st.nan_samples = h.nan_samples;
    st.iterations = h.iterations;
    st.kernel_launches = 1;
    if (ctx->profile_stages) st.shade_ms = ms;  // one kernel: all stages in one
  } catch (...) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    throw;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

// Accumulates samples [sample_begin, sample_begin+sample_count) of every pixel into d_accum.
// Work is enqueued on the lanes' own streams, which are ordered after everything already on `s`;
// `s` is ordered after the lanes when the call returns (it also blocks the host until then).
void wf_render(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count, uint64_t seed, float* d_accum,
               cudaStream_t s) {
  // Automatic choice: the register-resident kernel is one launch with no queue pools, the wavefront is faster
  // per path once its queues fill.  Measured on Cornell (tools/gpu_small_frames.py, device ms wavefront /
  // resident): 0.07 M paths 0.44 / 0.13, 0.26 M 0.57 / 0.27, 1.05 M 0.92 / 0.92, 4.2 M 2.4 / 3.4.
  constexpr uint64_t kResidentBelowPaths = 768u << 10;
  if (use_small(ctx) && (ctx->pipeline_cfg == 2 || (ctx->pipeline_cfg == 0 &&
                                                       (uint64_t)ctx->flat.raygen.width * ctx->flat.raygen.height * sample_count < kResidentBelowPaths)))
    return resident_render(ctx, sample_begin, sample_count, seed, d_accum, s);
  const FlatScene& f = ctx->flat;
  const uint64_t npix = (uint64_t)f.raygen.width * f.raygen.height;
  const uint64_t total = npix * sample_count;
  ctx->stats = TutuRenderStats{};
  if (total == 0) return;

  // lanes: split the samples; a lane never gets less than ~one wavefront of paths
  // Defaults measured on Cornell 1024^2 (tools/gpu_lanes.py, Mpaths/s): 1 lane x 4 / 8 / 16 / 32 Mi paths in
  // flight = 1556 / 1634 / 1667 / 1680.  Two lanes with every kernel launched at its FULL persistent grid
  // (the lanes then mostly alternate; one lane's next kernel fills the SMs that the other's draining kernel
  // frees): 2 x 16 Mi = 1725 vs 1659 for 1 x 16 Mi.  (Grids halved per lane: 1441.)
  // Defaults (Cornell 1024^2, tools/gpu_cornell_lanes.py): scenes shaded in queue order run one lane of 32 Mi paths —
  // wf_shade's 64-thread blocks leave no tails for a second lane to fill, and longer launches amortise the per-iteration
  // overhead (1 x 16 / 1 x 32 / 1 x 64 / 2 x 16 Mi: 2155 / 2185 / 2200 / 2120 Mpaths/s); scenes shaded through class
  // lists (256-thread blocks) keep two lanes of 16 Mi (glass scene: 934 against 932 Mpaths/s for 1 x 32 Mi).
  const bool one_lane = !ctx->sort_by_class;
  int n_lanes = ctx->lanes_cfg > 0 ? ctx->lanes_cfg : (one_lane ? 1 : 2);
  const uint64_t cap_cfg = ctx->paths_in_flight_cfg ? ctx->paths_in_flight_cfg : (uint64_t)(one_lane ? 32 : 16) << 20;
  while (n_lanes > 1 && (sample_count < (uint32_t)n_lanes || total / n_lanes < cap_cfg / 2)) --n_lanes;
  if ((int)ctx->wf_lanes.size() < n_lanes) ctx->wf_lanes.resize(n_lanes);
  // the cached grids depend on the scene through the kernel variants and the traversal-stack size
  const bool small = use_small(ctx);
  const bool wide = !small && ctx->dev.wide != nullptr;
  const bool stack_shared = ctx->stack_shared;
  const size_t want_stack = small ? 0 : (wide ? wide_stack_smem(ctx, 256) : (stack_shared ? stack_smem(ctx, 256, false) : 0));
  const size_t want_stack_any = (small || wide || !stack_shared) ? 0 : stack_smem(ctx, 256, true);
  if (ctx->grid_lanes != n_lanes || ctx->grid_small != small || ctx->grid_shade_block != ctx->shade_block ||
      ctx->grid_stack_smem != want_stack || ctx->grid_wide != wide || ctx->grid_stack_shared != stack_shared) {
    ctx->grid_stack_smem = want_stack;
    ctx->grid_stack_shared = stack_shared;
    ctx->grid_wide = wide;
    ctx->grid_shade_block = ctx->shade_block;
    ctx->grid_small = small;
    int div = 1;
#ifdef TUTU_EXPERIMENTS
    if (getenv("TUTU_GRID_SPLIT")) div = n_lanes;  // split the resident blocks between the lanes
#endif
    auto sized = [&](int full) { return ctx->sm_count * std::max(1, full / ctx->sm_count / div); };
    ctx->grid_extend = sized(small  ? persistent_grid(ctx, wf_extend_small, kSmallBlock)
                             : wide ? wide_grids_for(ctx).wf_extend
                             : stack_shared ? persistent_grid(ctx, wf_extend<0>, 256, want_stack)
                                            : persistent_grid(ctx, wf_extend<2>, 256));
    ctx->grid_shade = sized(persistent_grid(ctx, wf_shade, ctx->shade_block, shade_stage_smem(ctx)));
    ctx->grid_shadow = sized(small  ? persistent_grid(ctx, wf_shadow_small, kSmallBlock)
                             : wide ? wide_grids_for(ctx).wf_shadow
                             : stack_shared ? persistent_grid(ctx, wf_shadow<0>, 256, want_stack_any)
                                            : persistent_grid(ctx, wf_shadow<2>, 256));
    ctx->grid_raygen = sized(persistent_grid(ctx, wf_raygen, 256));
    ctx->grid_lanes = n_lanes;
  }
  RayGenK rk;
  fill_raygen(f, &rk);

  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  std::vector<StageTimer> timers;
  struct Run {
    uint32_t s_begin, s_count;
    int cur = 0;
    bool done = false;
  };
  std::vector<Run> runs(n_lanes);
  struct BlockGraph {  // one captured block of iterations per lane, valid for this render only (seed, sample range)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;
    BlockGraph() = default;
    BlockGraph(const BlockGraph&) = delete;
    BlockGraph& operator=(const BlockGraph&) = delete;
    ~BlockGraph() {
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
    }
  };
  std::vector<BlockGraph> graphs(n_lanes);
  uint64_t launches = 0;
  try {
    for (int k = 0; k < n_lanes; ++k) {  // queue pools first: a (re)allocation is not device time of the render
      const uint32_t b0 = (uint32_t)((uint64_t)sample_count * k / n_lanes);
      const uint32_t b1 = (uint32_t)((uint64_t)sample_count * (k + 1) / n_lanes);
      runs[k].s_begin = sample_begin + b0;
      runs[k].s_count = b1 - b0;
      lane_prepare(ctx, ctx->wf_lanes[k], std::min<uint64_t>(cap_cfg, std::max<uint64_t>(npix * runs[k].s_count, 1)));
    }
    CUDA_TRY(cudaEventRecord(e0, s));
    for (int k = 0; k < n_lanes; ++k) {
      WfLane& L = ctx->wf_lanes[k];
      const uint64_t lane_total = npix * runs[k].s_count;
      L.b.accum = d_accum;
      CUDA_TRY(cudaStreamWaitEvent(L.stream, e0, 0));
      WfCtl h{};
      h.total_paths = lane_total;
      *L.ctl_host = h;
      CUDA_TRY(cudaMemcpyAsync(L.b.ctl, L.ctl_host, sizeof(WfCtl), cudaMemcpyHostToDevice, L.stream));
      timers.emplace_back(ctx->profile_stages != 0);
      runs[k].done = lane_total == 0;
    }
    // One wavefront iteration of lane k on its stream: raygen -> extend -> (classify) -> shade -> shadow + the two
    // one-thread control kernels.  Returns the number of launches.
    auto enqueue_iteration = [&](int k, bool first_iteration) -> uint64_t {
      WfLane& L = ctx->wf_lanes[k];
      cudaStream_t ls = L.stream;
      StageTimer& timer = timers[k];
      const int cur = runs[k].cur;
      uint64_t n_launch = 6;
      timer.mark(0, ls);
      wf_raygen<<<ctx->grid_raygen, 256, 0, ls>>>(L.b, cur, rk, runs[k].s_begin);
      wf_ctl_after_raygen<<<1, 1, 0, ls>>>(L.b.ctl, L.b.capacity);
      timer.mark(1, ls);
      if (small)
        wf_extend_small<<<ctx->grid_extend, kSmallBlock, 0, ls>>>(ctx->dev, ctx->small, L.b, cur);
      else if (wide)
        CUDA_TRY(wide_launch_wf_extend(ctx->grid_extend, want_stack, ls, ctx->dev, L.b, cur));
      else if (stack_shared)
        wf_extend<0><<<ctx->grid_extend, 256, want_stack, ls>>>(ctx->dev, ctx->small, L.b, cur);
      else
        wf_extend<2><<<ctx->grid_extend, 256, 0, ls>>>(ctx->dev, ctx->small, L.b, cur);
      if (first_iteration && k + 1 < n_lanes) {
        // stagger the lanes by one stage so that unlike kernels (traverse / shade) overlap
        CUDA_TRY(cudaEventRecord(L.ev_done, ls));
        CUDA_TRY(cudaStreamWaitEvent(ctx->wf_lanes[k + 1].stream, L.ev_done, 0));
      }
      timer.mark(2, ls);
      if (L.b.class_perm) {
        wf_classify<<<ctx->sm_count * 8, 256, 0, ls>>>(ctx->dev, L.b);
        n_launch += 1;
      }
      wf_shade<<<ctx->grid_shade, ctx->shade_block, shade_stage_smem(ctx), ls>>>(ctx->dev, L.pool_map, L.b, cur, seed);
      timer.mark(3, ls);
      if (small)
        wf_shadow_small<<<ctx->grid_shadow, kSmallBlock, 0, ls>>>(ctx->dev, ctx->small, L.b, cur ^ 1);
      else if (wide)
        CUDA_TRY(wide_launch_wf_shadow(ctx->grid_shadow, want_stack, ls, ctx->dev, L.b, cur ^ 1));
      else if (stack_shared)
        wf_shadow<0><<<ctx->grid_shadow, 256, want_stack_any, ls>>>(ctx->dev, ctx->small, L.b, cur ^ 1);
      else
        wf_shadow<2><<<ctx->grid_shadow, 256, 0, ls>>>(ctx->dev, ctx->small, L.b, cur ^ 1);
      timer.mark(0, ls);
      wf_ctl_after_iter<<<1, 1, 0, ls>>>(L.b.ctl);
      runs[k].cur ^= 1;
      return n_launch;
    };
    // The host looks at a lane's control block every kBlockIters iterations.  The first block is enqueued launch
    // by launch (it carries the stagger between the lanes); every further block of a lane is ONE graph launch: the
    // block's launches + the control-block read-back, captured once per render from the lane's stream (an even
    // number of iterations, so the ping-pong index `cur` is the same at every block start).
    constexpr int kBlockIters = 8;
    static_assert(kBlockIters % 2 == 0, "a block must leave the ping-pong index where it found it");
    const bool use_graphs = ctx->profile_stages == 0;
    for (uint64_t block = 0;; ++block) {
      bool any = false;
      for (int k = 0; k < n_lanes; ++k) {
        if (runs[k].done) continue;
        any = true;
        WfLane& L = ctx->wf_lanes[k];
        if (block == 0 || !use_graphs) {
          for (int it = 0; it < kBlockIters; ++it) launches += enqueue_iteration(k, block == 0 && it == 0);
          CUDA_TRY(cudaMemcpyAsync(L.ctl_host, L.b.ctl, sizeof(WfCtl), cudaMemcpyDeviceToHost, L.stream));
          continue;
        }
        if (!graphs[k].exec) {
          CUDA_TRY(cudaStreamBeginCapture(L.stream, cudaStreamCaptureModeThreadLocal));
          uint64_t n_launch = 0;
          try {
            for (int it = 0; it < kBlockIters; ++it) n_launch += enqueue_iteration(k, false);
            CUDA_TRY(cudaMemcpyAsync(L.ctl_host, L.b.ctl, sizeof(WfCtl), cudaMemcpyDeviceToHost, L.stream));
          } catch (...) {
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(L.stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            throw;
          }
          CUDA_TRY(cudaStreamEndCapture(L.stream, &graphs[k].graph));
          CUDA_TRY(cudaGraphInstantiate(&graphs[k].exec, graphs[k].graph, 0));
          graphs[k].launches = n_launch;
        }
        CUDA_TRY(cudaGraphLaunch(graphs[k].exec, L.stream));
        launches += graphs[k].launches;
      }
      if (!any) break;
      CUDA_TRY(cudaGetLastError());
      for (int k = 0; k < n_lanes; ++k) {
        if (runs[k].done) continue;
        WfLane& L = ctx->wf_lanes[k];
        CUDA_TRY(cudaStreamSynchronize(L.stream));
        if (L.ctl_host->done) {
          runs[k].done = true;
          timers[k].mark(0, L.stream);
        }
      }
    }
    for (int k = 0; k < n_lanes; ++k) {
      WfLane& L = ctx->wf_lanes[k];
      CUDA_TRY(cudaEventRecord(L.ev_done, L.stream));
      CUDA_TRY(cudaStreamWaitEvent(s, L.ev_done, 0));
    }
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    TutuRenderStats& st = ctx->stats;
    st.gpu_ms = ms;
    st.paths = total;
    for (int k = 0; k < n_lanes; ++k) {
      const WfCtl& c = *ctx->wf_lanes[k].ctl_host;
      st.extend_rays += c.sum_extend;
      st.shadow_rays += c.sum_shadow;
      st.nan_samples += c.nan_samples;
      st.iterations += c.iterations;
      timers[k].resolve(&st);
    }
    st.shade_calls = st.extend_rays;
    st.kernel_launches = launches;
  } catch (...) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    throw;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

// ---------------------------------------------------------------------------------------------
// BDPT host loop (bdpt.cuh)
// ---------------------------------------------------------------------------------------------
void bdpt_prepare(TutuCtx* ctx, BdptLane& L, uint64_t cap) {
  cap = (cap + 255) & ~(uint64_t)255;
  {
    // float4 arrays: 6 vertex arrays x 14 slots, walk queue 2 x 3 x (2 cap), hit (2 cap), shadow 3 x (7 cap)
    const size_t n_f4 = (size_t)6 * kBdptSlots + 2 * 3 * 2 + 2 + 3 * kBdptMaxLen;
    const size_t bytes = n_f4 * cap * sizeof(float4) + (size_t)kBdptSlots * cap * sizeof(float) + 2 * cap;
    L.pool.ensure(bytes);  // grows only; the arrays are laid out for THIS batch size
  }
  BdptBuffers& b = L.b;
  float4* p = L.pool.as<float4>();
  b.vP = p, p += kBdptSlots * cap;
  b.vNg = p, p += kBdptSlots * cap;
  b.vNs = p, p += kBdptSlots * cap;
  b.vT = p, p += kBdptSlots * cap;
  b.vM = p, p += kBdptSlots * cap;
  b.vMis = p, p += kBdptSlots * cap;
  for (int k = 0; k < 2; ++k) {
    b.q_o[k] = p, p += 2 * cap;
    b.q_d[k] = p, p += 2 * cap;
    b.q_tp[k] = p, p += 2 * cap;
  }
  b.hit = p, p += 2 * cap;
  b.sh_o = p, p += kBdptMaxLen * cap;
  b.sh_d = p, p += kBdptMaxLen * cap;
  b.sh_c = p, p += kBdptMaxLen * cap;
  b.vMet = reinterpret_cast<float*>(p);
  b.nE = reinterpret_cast<unsigned char*>(b.vMet + kBdptSlots * cap);
  b.nL = b.nE + cap;
  L.ctl.ensure(sizeof(BdptCtl));
  b.ctl = L.ctl.as<BdptCtl>();
  b.cap = (unsigned)cap;
  if (!L.ctl_host) CUDA_TRY(cudaMallocHost(&L.ctl_host, sizeof(BdptCtl)));
  if (!L.stream) CUDA_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
  if (!L.ev_done) CUDA_TRY(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
}

// Adds the strategy sums of samples [sample_begin, sample_begin+sample_count) of every pixel into
// d_accum.  Batches alternate between two lanes (own pools, own streams): the late walk iterations and
// the long path lengths of a batch carry few rays and leave most SMs idle, which the other lane's
// batch fills.  The lanes' streams are ordered after everything already on `s`, and `s` after them;
// the call returns once the work is done (stats readback).
void bdpt_render(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count, uint64_t seed, float* d_accum,
                 cudaStream_t s) {
  const FlatScene& f = ctx->flat;
  const uint64_t npix = (uint64_t)f.raygen.width * f.raygen.height;
  const uint64_t total = npix * sample_count;
  ctx->stats = TutuRenderStats{};
  if (total == 0) return;
  // samples per batch (tools/gpu_bdpt_batch.py, Veach 800x600, Msamples/s, one lane): 0.5 / 1 / 2 / 4 / 8 Mi =
  // 45.9 / 51.1 / 54.2 / 56.2 / 57.3; 4 Mi samples hold 8.2 GB of vertices and queues
  const uint64_t cap_cfg = ctx->paths_in_flight_cfg ? ctx->paths_in_flight_cfg : (uint64_t)4 << 20;
  const uint64_t cap = std::min<uint64_t>(cap_cfg, total);
  const uint64_t n_batches = (total + cap - 1) / cap;
  int n_lanes = ctx->lanes_cfg > 0 ? std::min(ctx->lanes_cfg, 2) : 2;
  if (n_batches < 2) n_lanes = 1;
  for (int k = 0; k < n_lanes; ++k) bdpt_prepare(ctx, ctx->bdpt_lanes[k], cap);
  const bool small = use_small(ctx);
  const BdptCam& cam = f.bdpt_cam;
  const int g_start = persistent_grid(ctx, bdpt_start, 256);
  const int g_vertex = persistent_grid(ctx, bdpt_vertex, 256);
  const int g_connect = persistent_grid(ctx, bdpt_connect, 256);
  const bool wide = !small && ctx->dev.wide != nullptr;
  const bool stack_shared = ctx->stack_shared;
  const size_t sm_stack = wide ? wide_stack_smem(ctx, 256) : (stack_shared ? stack_smem(ctx, 256, false) : 0);
  const size_t sm_stack_any = (wide || !stack_shared) ? 0 : stack_smem(ctx, 256, true);
  const int g_extend = small  ? persistent_grid(ctx, q_extend<1>, 256)
                       : wide ? wide_grids_for(ctx).q_extend
                      : stack_shared ? persistent_grid(ctx, q_extend<0>, 256, sm_stack)
                                     : persistent_grid(ctx, q_extend<2>, 256);
  const int g_shadow = small  ? persistent_grid(ctx, q_shadow_add<1>, 256)
                       : wide ? wide_grids_for(ctx).q_shadow
                      : stack_shared ? persistent_grid(ctx, q_shadow_add<0>, 256, sm_stack_any)
                                     : persistent_grid(ctx, q_shadow_add<2>, 256);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  uint64_t launches = 0;
  try {
    CUDA_TRY(cudaEventRecord(e0, s));
    for (int k = 0; k < n_lanes; ++k) {
      BdptLane& L = ctx->bdpt_lanes[k];
      L.b.accum = d_accum;
      CUDA_TRY(cudaStreamWaitEvent(L.stream, e0, 0));
      CUDA_TRY(cudaMemsetAsync(L.b.ctl, 0, sizeof(BdptCtl), L.stream));
    }
    // queue tracer of this render (see TutuCtx::bdpt_tracer_cfg): forced, already measured for this scene, or measured now
    const bool tree_kernels = !small && !wide;
    int tracer = ctx->bdpt_tracer_cfg ? ctx->bdpt_tracer_cfg - 1 : (ctx->bdpt_tracer_tuned >= 0 ? ctx->bdpt_tracer_tuned : 0);
    const bool tune = tree_kernels && !ctx->bdpt_tracer_cfg && ctx->bdpt_tracer_tuned < 0 && n_batches >= kBdptTuneMinBatches;
    struct TuneEvents {  // destroyed on every way out of the render
      cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
      ~TuneEvents() {
        for (cudaEvent_t x : e)
          if (x) cudaEventDestroy(x);
      }
    } tune_events;
    cudaEvent_t* et = tune_events.e;
    if (tune)
      for (int k = 0; k < 3; ++k) CUDA_TRY(cudaEventCreate(&et[k]));
    DevScene dv = ctx->dev;
    uint64_t batch = 0;
    for (uint64_t first = 0; first < total; first += cap, ++batch) {
      // while measuring, batches 0 (lanes) and 1 (packets) run alone, back to back on the first lane's stream; the
      // lanes go first, so whatever a cold start costs counts against them
      const bool measuring = tune && batch < 2;
      BdptLane& L = ctx->bdpt_lanes[measuring ? 0 : batch % n_lanes];
      BdptBuffers& b = L.b;
      cudaStream_t ls = L.stream;
      dv.queue_lanes = tree_kernels ? (measuring ? 1 - (int)batch : tracer) : 0;
      if (measuring) CUDA_TRY(cudaEventRecord(et[batch], ls));
      const unsigned n = (unsigned)std::min<uint64_t>(cap, total - first);
      bdpt_start<<<g_start, 256, 0, ls>>>(ctx->dev, cam, b, first, n, sample_begin, seed);
      bdpt_ctl_begin<<<1, 1, 0, ls>>>(b.ctl, 2 * n);
      launches += 2;
      int cur = 0;
      for (int it = 0; it < kBdptMaxLen; ++it) {  // vertices it+1 of both walks
        if (small)
          q_extend<1><<<g_extend, 256, 0, ls>>>(ctx->dev, ctx->small, b.q_o[cur], b.q_d[cur], b.hit, &b.ctl->n_cur, &b.ctl->cursor_extend);
        else if (wide)
          CUDA_TRY(wide_launch_q_extend(g_extend, sm_stack, ls, ctx->dev, b.q_o[cur], b.q_d[cur], b.hit, &b.ctl->n_cur, &b.ctl->cursor_extend));
        else if (stack_shared)
          q_extend<0><<<g_extend, 256, sm_stack, ls>>>(dv, ctx->small, b.q_o[cur], b.q_d[cur], b.hit, &b.ctl->n_cur, &b.ctl->cursor_extend);
        else
          q_extend<2><<<g_extend, 256, 0, ls>>>(dv, ctx->small, b.q_o[cur], b.q_d[cur], b.hit, &b.ctl->n_cur, &b.ctl->cursor_extend);
        bdpt_vertex<<<g_vertex, 256, 0, ls>>>(ctx->dev, cam, b, cur, first, sample_begin, seed);
        bdpt_ctl_after_walk<<<1, 1, 0, ls>>>(b.ctl);
        launches += 3;
        cur ^= 1;
      }
      for (int len = 1; len <= kBdptMaxLen; ++len) {
        bdpt_connect<<<g_connect, 256, 0, ls>>>(ctx->dev, cam, b, len, first, n);
        if (small)
          q_shadow_add<1><<<g_shadow, 256, 0, ls>>>(ctx->dev, ctx->small, b.sh_o, b.sh_d, b.sh_c, b.accum, &b.ctl->n_shadow, &b.ctl->cursor_shadow);
        else if (wide)
          CUDA_TRY(wide_launch_q_shadow_add(g_shadow, sm_stack, ls, ctx->dev, b.sh_o, b.sh_d, b.sh_c, b.accum, &b.ctl->n_shadow, &b.ctl->cursor_shadow));
        else if (stack_shared)
          q_shadow_add<0><<<g_shadow, 256, sm_stack_any, ls>>>(dv, ctx->small, b.sh_o, b.sh_d, b.sh_c, b.accum, &b.ctl->n_shadow, &b.ctl->cursor_shadow);
        else
          q_shadow_add<2><<<g_shadow, 256, 0, ls>>>(dv, ctx->small, b.sh_o, b.sh_d, b.sh_c, b.accum, &b.ctl->n_shadow, &b.ctl->cursor_shadow);
        bdpt_ctl_after_shadow<<<1, 1, 0, ls>>>(b.ctl);
        launches += 3;
      }
      CUDA_TRY(cudaGetLastError());
      if (measuring) {
        CUDA_TRY(cudaEventRecord(et[batch + 1], ls));
        if (batch == 1) {  // both measured: the rest of the render (and later renders of this scene) take the faster one
          CUDA_TRY(cudaEventSynchronize(et[2]));
          CUDA_TRY(cudaEventElapsedTime(&ctx->bdpt_tune_ms[1], et[0], et[1]));
          CUDA_TRY(cudaEventElapsedTime(&ctx->bdpt_tune_ms[0], et[1], et[2]));
          tracer = ctx->bdpt_tune_ms[1] < ctx->bdpt_tune_ms[0] ? 1 : 0;
          ctx->bdpt_tracer_tuned = tracer;
        }
      }
    }
    for (int k = 0; k < n_lanes; ++k) {
      BdptLane& L = ctx->bdpt_lanes[k];
      CUDA_TRY(cudaMemcpyAsync(L.ctl_host, L.b.ctl, sizeof(BdptCtl), cudaMemcpyDeviceToHost, L.stream));
      CUDA_TRY(cudaEventRecord(L.ev_done, L.stream));
      CUDA_TRY(cudaStreamWaitEvent(s, L.ev_done, 0));
    }
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    TutuRenderStats& st = ctx->stats;
    st.gpu_ms = ms;
    st.paths = total;
    for (int k = 0; k < n_lanes; ++k) {
      st.extend_rays += ctx->bdpt_lanes[k].ctl_host->sum_extend;
      st.shadow_rays += ctx->bdpt_lanes[k].ctl_host->sum_shadow;
    }
    st.shade_calls = 0;
    st.kernel_launches = launches;
    st.iterations = n_batches;
  } catch (...) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    throw;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
// Guard bands of every live device allocation (only in a -DTUTU_GUARDS build, see DevBuf).
extern "C" int tutu_debug_guard_check(uint64_t* n_buffers, uint64_t* n_bad_bytes) {
  if (!n_buffers || !n_bad_bytes) return fail(nullptr, TUTU_E_INVALID, "tutu_debug_guard_check: null argument");
  *n_buffers = 0;
  *n_bad_bytes = 0;
#ifdef TUTU_GUARDS
  try {
    CUDA_TRY(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_guard_mu);
    *n_bad_bytes = g_guard_bad_freed;
    for (const auto& b : g_guard_live) {
      *n_bad_bytes += guard_bad_bytes(b.first, b.second);
      ++*n_buffers;
    }
    return TUTU_OK;
  } catch (const CudaError& e) {
    return fail_cuda(nullptr, e);
  } catch (...) {
    return fail(nullptr, TUTU_E_INVALID, "tutu_debug_guard_check: unexpected exception");
  }
#else
  return fail(nullptr, TUTU_E_STATE, "tutu_debug_guard_check: this library was built without -DTUTU_GUARDS");
#endif
}

// Self-test of the detector: overwrites `n_bytes` of the upper band of the oldest live allocation.
extern "C" int tutu_debug_guard_poke(uint32_t n_bytes) {
#ifdef TUTU_GUARDS
  std::lock_guard<std::mutex> lock(g_guard_mu);
  if (g_guard_live.empty() || n_bytes > kGuardBytes) return fail(nullptr, TUTU_E_STATE, "tutu_debug_guard_poke: nothing to poke");
  const auto& b = g_guard_live.front();
  if (cudaMemset(static_cast<char*>(b.first) + b.second, 0, n_bytes) != cudaSuccess) return fail(nullptr, TUTU_E_CUDA, "tutu_debug_guard_poke: cudaMemset failed");
  return TUTU_OK;
#else
  (void)n_bytes;
  return fail(nullptr, TUTU_E_STATE, "tutu_debug_guard_poke: this library was built without -DTUTU_GUARDS");
#endif
}

extern "C" int tutu_ctx_create(int device, TutuCtx** out) {
  if (!out) return fail(nullptr, TUTU_E_INVALID, "tutu_ctx_create: null out pointer");
  *out = nullptr;
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      cudaGetLastError();
      return fail(nullptr, TUTU_E_CUDA,
                  std::string("tutu_ctx_create: no usable CUDA device (") +
                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                      "); libtutu_b200 has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(nullptr, TUTU_E_INVALID, "tutu_ctx_create: bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      return fail(nullptr, TUTU_E_CUDA,
                  std::string("tutu_ctx_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                      std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
    std::unique_ptr<TutuCtx> ctx(new TutuCtx());
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    *out = ctx.release();
    return TUTU_OK;
  } catch (const CudaError& e) {
    return fail_cuda(nullptr, e);
  } catch (const std::bad_alloc&) {
    return fail(nullptr, TUTU_E_NOMEM, "out of host memory");
  } catch (const std::exception& e) {
    return fail(nullptr, TUTU_E_INVALID, std::string("tutu_ctx_create: ") + e.what());
  } catch (...) {
    return fail(nullptr, TUTU_E_INVALID, "tutu_ctx_create: unexpected exception");
  }
}

extern "C" void tutu_ctx_destroy(TutuCtx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  for (cudaStream_t st : ctx->slot_streams)
    if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : ctx->slot_last_use)
    if (ev) cudaEventDestroy(ev);
  delete ctx;
}

extern "C" int tutu_scene_upload(TutuCtx* ctx, const TutuSceneDesc* desc) {
  API_BEGIN(ctx)
  using Clock = std::chrono::steady_clock;
  auto ms_since = [](Clock::time_point t) { return std::chrono::duration<float, std::milli>(Clock::now() - t).count(); };
  const Clock::time_point t_start = Clock::now();
  // TUTU_BUILD_AUTO: the binned-SAH tree, built on the device for scenes of more than 64 primitives (the same tree as the
  // host builder's, 9.6 instead of 163 ms for 999 698 triangles; tools/gpu_build_ab.py) and on the host below that.  The
  // two other device builders are faster still to build but cost 2-2.7x the node visits per ray.
  const bool want_ploc = desc && ctx->builder_cfg == TUTU_BUILD_DEVICE_PLOC;
  const bool want_sah = desc && (ctx->builder_cfg == TUTU_BUILD_DEVICE_SAH || (ctx->builder_cfg == TUTU_BUILD_AUTO && desc->n_prims > 64u));
  const bool want_device = desc && (ctx->builder_cfg == TUTU_BUILD_DEVICE_LBVH || want_ploc || want_sah);
  FlatScene fs;
  int rc = flatten_scene(desc, &fs, false);  // the traversal tree is built (and timed) below
  if (rc != TUTU_OK) return fail(ctx, rc, get_error());
  TutuUploadStats us{};
  us.builder = TUTU_BUILD_HOST_SAH;
  us.flatten_ms = ms_since(t_start);
  cudaStream_t s = ctx->stream;
  bool device_tree = false;
  const Clock::time_point t_build = Clock::now();
  if (!want_device) {
    build_host_fast_tree(&fs);
  } else {
    bool finite = fs.n_prims >= 3 && fs.n_prims <= (1u << 28);  // device_bvh.cu indexes leaf ranges with int arithmetic
    for (size_t k = 0; finite && k < fs.leaf_box.size(); ++k)
      for (int a = 0; a < 3; ++a) finite = finite && std::isfinite(fs.leaf_box[k].lo[a]) && std::isfinite(fs.leaf_box[k].hi[a]);
    if (finite) {
      upload_vec(ctx->d_leaf_box, fs.leaf_box, s);
      upload_vec(ctx->d_leaf_code, fs.leaf_code, s);
      ctx->d_inner_fast.ensure((size_t)(fs.n_prims - 1) * sizeof(InnerNode));
      ctx->d_build_scratch.ensure(want_sah    ? device_build_sah_scratch_bytes(fs.n_prims)
                                  : want_ploc ? device_build_ploc_scratch_bytes(fs.n_prims)
                                              : device_build_lbvh_scratch_bytes(fs.n_prims));
      uint32_t depth = 0;
      if (want_sah) {
        int lg = 0;  // the host builder's precondition (host_scene.cpp: build_fast_tree)
        while ((1ull << lg) < fs.n_prims) ++lg;
        if (lg + 2 <= kFastTreeMaxDepth)
          CUDA_TRY(device_build_sah(ctx->d_leaf_box.as<float>(), ctx->d_leaf_code.as<uint32_t>(), fs.n_prims, kFastTreeMaxDepth,
                                    ctx->d_inner_fast.p, &depth, nullptr, ctx->sm_count, ctx->d_build_scratch.p, s));
      } else if (want_ploc)
        CUDA_TRY(device_build_ploc(ctx->d_leaf_box.as<float>(), ctx->d_leaf_code.as<uint32_t>(), fs.n_prims, fs.root_box.lo,
                                   fs.root_box.hi, ctx->d_inner_fast.p, &depth, nullptr, ctx->sm_count, ctx->d_build_scratch.p, s));
      else
        CUDA_TRY(device_build_lbvh(ctx->d_leaf_box.as<float>(), ctx->d_leaf_code.as<uint32_t>(), fs.n_prims, fs.root_box.lo,
                                   fs.root_box.hi, ctx->d_inner_fast.p, &depth, ctx->sm_count, ctx->d_build_scratch.p, s));
      if (depth >= 1 && depth <= (uint32_t)kFastTreeMaxDepth) {
        device_tree = true;
        fs.root_ref_fast = 0;
        fs.depth_fast = depth;
        us.builder = want_sah ? TUTU_BUILD_DEVICE_SAH : want_ploc ? TUTU_BUILD_DEVICE_PLOC : TUTU_BUILD_DEVICE_LBVH;
      }
    }
    if (!device_tree) build_host_fast_tree(&fs);  // non-finite boxes, tiny scenes, or a tree deeper than the stacks
  }
  us.tree_build_ms = ms_since(t_build);
  if (ctx->use_wide) {
    if (device_tree) {  // the wide collapse runs on the host: fetch the binary tree it collapses
      fs.inner_fast.resize(fs.n_prims - 1);
      CUDA_TRY(cudaMemcpyAsync(fs.inner_fast.data(), ctx->d_inner_fast.p, fs.inner_fast.size() * sizeof(InnerNode), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
    }
    build_wide_tree(&fs);
  }
  if (std::max(fs.depth, fs.depth_fast) > (uint32_t)kStackSize)
    return fail(ctx, TUTU_E_INVALID, "scene: BVH deeper than the traversal stack (" + std::to_string(fs.depth) + ")");
  const Clock::time_point t_h2d = Clock::now();
  upload_vec(ctx->d_inner, fs.inner, s);
  if (!device_tree) upload_vec(ctx->d_inner_fast, fs.inner_fast, s);
  upload_vec(ctx->d_wide, fs.wide, s);
  upload_vec(ctx->d_wleaf, fs.wleaf, s);
  upload_vec(ctx->d_wbox, fs.wbox, s);
  upload_vec(ctx->d_geom, fs.geom, s);
  upload_vec(ctx->d_shade, fs.shade, s);
  upload_vec(ctx->d_leaftex, fs.leaftex, s);
  upload_vec(ctx->d_slot_to_prim, fs.slot_to_prim, s);
  upload_vec(ctx->d_materials, fs.materials, s);
  upload_vec(ctx->d_lights, fs.lights, s);
  upload_vec(ctx->d_texels, fs.texels, s);
  for (int c = 0; c < 4; ++c) upload_vec(ctx->d_texh[c], fs.tex_headers[c], s);
  CUDA_TRY(cudaStreamSynchronize(s));
  us.h2d_ms = ms_since(t_h2d);
  us.tree_depth = fs.depth_fast;
  DevScene& d = ctx->dev;
  d.inner = ctx->d_inner.as<float4>();
  d.inner_fast = ctx->d_inner_fast.as<float4>();
  d.root_ref_fast = fs.root_ref_fast;
  d.wide = (fs.wide.empty() || !ctx->use_wide) ? nullptr : ctx->d_wide.as<float4>();
  d.wleaf = ctx->d_wleaf.as<float4>();
  d.wbox = ctx->d_wbox.as<float4>();
  d.geom = ctx->d_geom.as<float4>();
  d.shade = ctx->d_shade.as<float4>();
  d.leaftex = fs.leaftex.empty() ? nullptr : ctx->d_leaftex.as<int4>();
  d.slot_to_prim = ctx->d_slot_to_prim.as<int>();
  d.materials = ctx->d_materials.as<float4>();
  d.lights = ctx->d_lights.as<float4>();
  for (int c = 0; c < 4; ++c) d.tex_headers[c] = ctx->d_texh[c].as<int4>();
  d.texels = ctx->d_texels.as<float4>();
  memcpy(d.root_lo, fs.root_box.lo, 12);
  memcpy(d.root_hi, fs.root_box.hi, 12);
  d.root_ref = fs.root_ref;
  d.empty = fs.empty ? 1 : 0;
  d.n_lights = (int)fs.lights.size();
  memcpy(d.bkg, fs.bkgcolor, 12);
  d.eta = fs.eta;
  d.prune_rel = 1.0f / 1024.0f;
  d.prune_abs = fs.max_edge * (1.0f / 512.0f);
  d.sphere_mask = 0u;
  ctx->small = SmallScene{};
  if (!fs.empty && fs.n_prims <= (uint32_t)kSmallMax) {
    ctx->small.n = (int)fs.n_prims;
    SmallScene& sm = ctx->small;
    for (uint32_t k = 0; k < fs.n_prims; ++k) {
      float2 b[3];
      for (int a = 0; a < 3; ++a) b[a] = make_float2(fs.leaf_box[k].lo[a], fs.leaf_box[k].hi[a]);
      int j = 0;  // bitwise-identical bounds share one slab test
      while (j < sm.n_boxes && memcmp(sm.box[j], b, sizeof(b)) != 0) ++j;
      if (j == sm.n_boxes) {
        j = sm.n_boxes++;
        memcpy(sm.box[j], b, sizeof(b));
      }
      sm.slots[j] |= 1u << k;
      if (fs.shade[k].flags & SHADE_SPHERE_BIT) d.sphere_mask |= 1u << k;
    }
    for (int j = sm.n_boxes; j < kSmallMax; ++j) memcpy(sm.box[j], sm.box[0], sizeof(sm.box[0]));  // slots = 0
  }
  {
    bool simple = true;
    for (const DevMaterial& m : fs.materials) simple = simple && m.type == TUTU_MAT_LAMBERTIAN;
    for (const LeafShade& ls : fs.shade) simple = simple && !(ls.flags & TEX_ACTIVE_BIT);
    ctx->shade_block = simple ? kShadeBlockSimple : TUTU_SHADE_BLOCK;
    ctx->sort_by_class = !simple;
#ifdef TUTU_EXPERIMENTS
    if (getenv("TUTU_NO_CLASS_SORT")) ctx->sort_by_class = false;
    if (const char* e = getenv("TUTU_SHADE_BLOCK_RT")) ctx->shade_block = atoi(e);
#endif
  }
  d.refill_min = kLanesRefillMin;
  d.leaf_batch = kLeafBatch;
  d.queue_lanes = 0;  // set per batch by bdpt_render
  d.lanes_leaf_batch = kLanesLeafBatch;
  ctx->bdpt_tracer_tuned = -1;  // a new scene: bdpt_render measures both queue tracers again
  // stack flavour of the tree kernels: see TutuCtx::stack_shared (a device-built tree has n - 1 nodes and no host copy)
  ctx->tree_bytes = (2 * fs.inner.size()) * sizeof(InnerNode) + fs.geom.size() * sizeof(LeafGeom);
  ctx->stack_shared = ctx->stack_cfg ? ctx->stack_cfg == 1 : ctx->tree_bytes > kLocalStackMaxTreeBytes;
#ifdef TUTU_EXPERIMENTS  // never in the shipped library: the pruning slack is part of the parity argument (trace.cuh)
  if (const char* e = getenv("TUTU_LEAF_BATCH")) d.leaf_batch = atoi(e);
  if (const char* e = getenv("TUTU_LANES_LEAF_BATCH")) d.lanes_leaf_batch = atoi(e);
  if (const char* e = getenv("TUTU_REFILL_MIN")) d.refill_min = atoi(e);
  if (const char* e = getenv("TUTU_PRUNE_REL")) d.prune_rel = (float)atof(e);
  if (const char* e = getenv("TUTU_PRUNE_ABS")) d.prune_abs = (float)atof(e);
#endif
  ctx->scene_bytes = (fs.inner.size() + fs.inner_fast.size()) * sizeof(InnerNode) + fs.geom.size() * sizeof(LeafGeom) +
                     fs.wide.size() * sizeof(WideNode) + fs.wleaf.size() * sizeof(WideLeaf) + fs.wbox.size() * sizeof(WideLeafBox) +
                     fs.shade.size() * sizeof(LeafShade) + fs.leaftex.size() * sizeof(LeafTex) +
                     fs.slot_to_prim.size() * 4 + fs.materials.size() * sizeof(DevMaterial) +
                     fs.lights.size() * sizeof(DevLight) + fs.texels.size() * 4;
  ctx->flat = std::move(fs);
  ctx->device_tree = device_tree;
  ctx->has_scene = true;
  us.total_ms = ms_since(t_start);
  ctx->upload_stats = us;
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_scene_info(const TutuCtx* ctx, TutuSceneInfo* out) {
  if (!ctx || !out) return fail(nullptr, TUTU_E_INVALID, "tutu_scene_info: null argument");
  if (!ctx->has_scene) return fail(const_cast<TutuCtx*>(ctx), TUTU_E_STATE, "no scene uploaded");
  const FlatScene& f = ctx->flat;
  out->n_prims = f.n_prims;
  out->n_nodes = f.n_ref_nodes;
  out->n_inner = (uint32_t)f.inner.size();
  out->depth = f.depth;
  out->n_lights = (uint32_t)f.lights.size();
  out->n_materials = (uint32_t)f.materials.size();
  out->width = (uint32_t)f.raygen.width;
  out->height = (uint32_t)f.raygen.height;
  out->device_bytes = ctx->scene_bytes;
  const bool wide = ctx->dev.wide != nullptr;
  out->trav_nodes = wide ? (uint32_t)f.wide.size() : (ctx->device_tree ? f.n_prims - 1 : (uint32_t)f.inner_fast.size());
  out->trav_depth = wide ? f.wide_depth : f.depth_fast;
  out->trav_width = wide ? 8 : 2;
  out->trav_node_bytes = wide ? (uint32_t)sizeof(WideNode) : (uint32_t)sizeof(InnerNode);
  out->trav_leaf_bytes = wide ? (uint32_t)sizeof(WideLeaf) : (uint32_t)sizeof(LeafGeom);
  out->reserved = 0;
  return TUTU_OK;
}

extern "C" int tutu_scene_builder(TutuCtx* ctx, int builder) {
  if (!ctx || builder < TUTU_BUILD_AUTO || builder > TUTU_BUILD_DEVICE_SAH)
    return fail(ctx, TUTU_E_INVALID, "tutu_scene_builder: builder must be TUTU_BUILD_AUTO, _HOST_SAH or _DEVICE_LBVH");
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->builder_cfg = builder;
  return TUTU_OK;
}

extern "C" int tutu_upload_stats(const TutuCtx* ctx, TutuUploadStats* out) {
  if (!ctx || !out) return fail(nullptr, TUTU_E_INVALID, "tutu_upload_stats: null argument");
  if (!ctx->has_scene) return fail(const_cast<TutuCtx*>(ctx), TUTU_E_STATE, "no scene uploaded");
  *out = ctx->upload_stats;
  return TUTU_OK;
}

extern "C" int tutu_scene_set_camera(TutuCtx* ctx, const TutuCamera* cam) {
  if (!ctx || !cam) return fail(ctx, TUTU_E_INVALID, "tutu_scene_set_camera: null argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!ctx->has_scene) return fail(ctx, TUTU_E_STATE, "no scene uploaded");
  RayGen rg;
  int rc = compute_raygen(cam, &rg);
  if (rc != TUTU_OK) return fail(ctx, rc, get_error());
  BdptCamConsts bc;
  rc = compute_bdpt_cam(cam, &bc);
  if (rc != TUTU_OK) return fail(ctx, rc, get_error());
  ctx->flat.raygen = rg;
  ctx->flat.bdpt_cam = bc;
  ctx->flat.camera = *cam;
  return TUTU_OK;
}

extern "C" int tutu_bdpt_queue_tracer(TutuCtx* ctx, int tracer) {
  if (!ctx || tracer < 0 || tracer > 2) return fail(ctx, TUTU_E_INVALID, "tutu_bdpt_queue_tracer: bad argument (0, 1 or 2)");
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->bdpt_tracer_cfg = tracer;
  return TUTU_OK;
}

extern "C" int tutu_bdpt_queue_tracer_measured(const TutuCtx* ctx, int* tracer, float* ms_packets, float* ms_lanes) {
  if (!ctx || !tracer || !ms_packets || !ms_lanes) return fail(const_cast<TutuCtx*>(ctx), TUTU_E_INVALID, "tutu_bdpt_queue_tracer_measured: null argument");
  *tracer = ctx->bdpt_tracer_tuned;
  *ms_packets = ctx->bdpt_tune_ms[0];
  *ms_lanes = ctx->bdpt_tune_ms[1];
  return TUTU_OK;
}

extern "C" int tutu_traversal_stack(TutuCtx* ctx, int where) {
  if (!ctx || where < 0 || where > 2) return fail(ctx, TUTU_E_INVALID, "tutu_traversal_stack: bad argument (0, 1 or 2)");
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->stack_cfg = where;
  if (ctx->has_scene) ctx->stack_shared = where ? where == 1 : ctx->tree_bytes > kLocalStackMaxTreeBytes;
  return TUTU_OK;
}

extern "C" int tutu_set_traversal_mode(TutuCtx* ctx, int mode) {
#ifdef TUTU_EXPERIMENTS
  const bool experimental = mode == 2 || (mode >= 10 && mode <= 17);  // trace_variants.cuh
#else
  const bool experimental = false;
#endif
  if (!ctx || !(mode == 0 || mode == 1 || mode == 3 || mode == 4 || mode == 6 || experimental))
    return fail(ctx, TUTU_E_INVALID, "tutu_set_traversal_mode: bad argument (0, 1, 3, 4 or 6)");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (mode == 6) {
    // Regular rays walk the compressed 8-wide collapse of the SAH tree (wide.cuh).  Bit-identical hits; measured
    // SLOWER than the binary walk on a B200 (the exact plane decode makes a node visit ~210
    // instructions and these kernels are issue bound), so it is an option, not the default.
    try {
      CUDA_TRY(cudaSetDevice(ctx->device));
      ctx->use_wide = true;
      if (ctx->has_scene && ctx->flat.wide.empty()) {
        CUDA_TRY(cudaDeviceSynchronize());  // nothing may still read the buffers that are replaced below
        if (ctx->device_tree && ctx->flat.inner_fast.empty()) {  // the collapse runs on the host
          ctx->flat.inner_fast.resize(ctx->flat.n_prims - 1);
          CUDA_TRY(cudaMemcpy(ctx->flat.inner_fast.data(), ctx->d_inner_fast.p, ctx->flat.inner_fast.size() * sizeof(InnerNode),
                              cudaMemcpyDeviceToHost));
        }
        build_wide_tree(&ctx->flat);
        upload_vec(ctx->d_wide, ctx->flat.wide, ctx->stream);
        upload_vec(ctx->d_wleaf, ctx->flat.wleaf, ctx->stream);
        upload_vec(ctx->d_wbox, ctx->flat.wbox, ctx->stream);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->dev.wleaf = ctx->d_wleaf.as<float4>();
        ctx->dev.wbox = ctx->d_wbox.as<float4>();
      }
      if (ctx->has_scene) ctx->dev.wide = ctx->flat.wide.empty() ? nullptr : ctx->d_wide.as<float4>();
    } catch (const CudaError& e) {
      return fail_cuda(ctx, e);
    } catch (const std::exception& e) {
      return fail(ctx, TUTU_E_NOMEM, std::string("tutu_set_traversal_mode: ") + e.what());
    }
    ctx->ray_binning = 1;
    ctx->traversal_mode = 0;
    return TUTU_OK;
  }
  ctx->use_wide = false;
  ctx->dev.wide = nullptr;
  if (mode == 3) {  // production walk, caller's ray order (no binning)
    ctx->traversal_mode = 0;
    ctx->ray_binning = 0;
    return TUTU_OK;
  }
  ctx->ray_binning = 1;
  ctx->traversal_mode = mode;
  return TUTU_OK;
}

extern "C" int tutu_trace_closest_device(TutuCtx* ctx, const float* d_rays, uint64_t n_rays, TutuHit* d_hits_out,
                                         void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (n_rays && (!d_rays || !d_hits_out)) return fail(ctx, TUTU_E_INVALID, "tutu_trace_closest_device: null buffer");
  launch_closest(ctx, d_rays, n_rays, d_hits_out, stream ? (cudaStream_t)stream : ctx->stream);
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_trace_any_device(TutuCtx* ctx, const float* d_rays, uint64_t n_rays, uint8_t* d_blocked_out,
                                     void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (n_rays && (!d_rays || !d_blocked_out)) return fail(ctx, TUTU_E_INVALID, "tutu_trace_any_device: null buffer");
  launch_any(ctx, d_rays, n_rays, d_blocked_out, stream ? (cudaStream_t)stream : ctx->stream);
  return TUTU_OK;
  API_END(ctx)
}

// Host-buffer batches: chunks of kHostChunk rays rotate over kHostSlots streams, so that the H2D copy of
// one chunk, the walk of the previous one and the D2H copy of the one before overlap (PCIe is full duplex;
// with pinned host buffers the batch costs ~max(H2D, walk, D2H) instead of their sum).  Measured on 2^24
// rays vs the 999 698-triangle height-field (tools/gpu_e2e.py, Mrays/s): 2 slots x 2 Mi 1223, 2 x 1 Mi 1213,
// 3 x 2 Mi 1386, 3 x 1 Mi 1523, 4 x 1 Mi 1524 (the H2D copy alone allows ~1700).
constexpr uint64_t kHostChunk = 1ull << 20;
constexpr int kHostSlots = 3;

template <class Out, class Launch>
static void trace_host_pipelined(TutuCtx* ctx, const float* rays, uint64_t n_rays, Out* out, Launch launch) {
#ifdef TUTU_EXPERIMENTS
  static const int slots_cfg = getenv("TUTU_HOST_SLOTS") ? std::min(kHostSlotsMax, std::max(1, atoi(getenv("TUTU_HOST_SLOTS")))) : kHostSlots;
  static const uint64_t chunk_cfg = getenv("TUTU_HOST_CHUNK_LOG2") ? 1ull << atoi(getenv("TUTU_HOST_CHUNK_LOG2")) : kHostChunk;
#else
  const int slots_cfg = kHostSlots;
  const uint64_t chunk_cfg = kHostChunk;
#endif
  const uint64_t chunk = std::min<uint64_t>(chunk_cfg, n_rays);
  const int slots = (int)std::min<uint64_t>((uint64_t)slots_cfg, (n_rays + chunk - 1) / chunk);
  cudaStream_t st[kHostSlotsMax];
  st[0] = ctx->stream;
  for (int k = 1; k < slots; ++k) {
    if (!ctx->slot_streams[k]) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->slot_streams[k], cudaStreamNonBlocking));
    st[k] = ctx->slot_streams[k];
  }
  for (int k = 0; k < slots; ++k) {
    ctx->d_chunk_rays[k].ensure(chunk * TUTU_RAY_FLOATS * sizeof(float));
    ctx->d_chunk_out[k].ensure(chunk * sizeof(Out));
  }
  int k = 0;
  for (uint64_t first = 0; first < n_rays; first += chunk, k = (k + 1) % slots) {
    const uint64_t m = std::min<uint64_t>(chunk, n_rays - first);
    CUDA_TRY(cudaMemcpyAsync(ctx->d_chunk_rays[k].p, rays + first * TUTU_RAY_FLOATS, m * TUTU_RAY_FLOATS * sizeof(float),
                             cudaMemcpyHostToDevice, st[k]));
    launch(ctx->d_chunk_rays[k].as<float>(), m, ctx->d_chunk_out[k].as<Out>(), st[k], k);
    CUDA_TRY(cudaMemcpyAsync(out + first, ctx->d_chunk_out[k].p, m * sizeof(Out), cudaMemcpyDeviceToHost, st[k]));
  }
  for (int j = 0; j < slots; ++j) CUDA_TRY(cudaStreamSynchronize(st[j]));
}

extern "C" int tutu_trace_closest(TutuCtx* ctx, const float* rays, uint64_t n_rays, TutuHit* hits_out) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (n_rays == 0) return TUTU_OK;
  if (!rays || !hits_out) return fail(ctx, TUTU_E_INVALID, "tutu_trace_closest: null buffer");
  trace_host_pipelined<TutuHit>(ctx, rays, n_rays, hits_out,
                                [&](const float* d_rays, uint64_t m, TutuHit* d_out, cudaStream_t s, int slot) {
                                  launch_closest(ctx, d_rays, m, d_out, s, slot);
                                });
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_trace_any(TutuCtx* ctx, const float* rays, uint64_t n_rays, uint8_t* blocked_out) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (n_rays == 0) return TUTU_OK;
  if (!rays || !blocked_out) return fail(ctx, TUTU_E_INVALID, "tutu_trace_any: null buffer");
  trace_host_pipelined<uint8_t>(ctx, rays, n_rays, blocked_out,
                                [&](const float* d_rays, uint64_t m, uint8_t* d_out, cudaStream_t s, int slot) {
                                  launch_any(ctx, d_rays, m, d_out, s, slot);
                                });
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_trace_count_visits(TutuCtx* ctx, const float* d_rays, uint64_t n_rays, int any_hit,
                                       uint64_t* nodes_out, uint64_t* prims_out) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!nodes_out || !prims_out || (n_rays && !d_rays)) return fail(ctx, TUTU_E_INVALID, "tutu_trace_count_visits: null argument");
  cudaStream_t s = ctx->stream;
  ctx->d_counts.ensure(256);
  unsigned long long* c = ctx->d_counts.as<unsigned long long>();
  CUDA_TRY(cudaMemsetAsync(c, 0, 16, s));
  if (n_rays) {
    const float4* rays = reinterpret_cast<const float4*>(d_rays);
    const int grid = ctx->sm_count * 8;
    if (ctx->dev.wide) {  // visits of the tree the production walk descends (irregular rays are not counted)
      const WideGrids& g = wide_grids_for(ctx);
      CUDA_TRY(wide_launch_count(any_hit != 0, grid, g.smem, s, ctx->dev, rays, n_rays, c));
    } else if (any_hit)
      k_trace_count<true><<<grid, 256, 0, s>>>(ctx->dev, rays, n_rays, c);
    else
      k_trace_count<false><<<grid, 256, 0, s>>>(ctx->dev, rays, n_rays, c);
    CUDA_TRY(cudaGetLastError());
  }
  unsigned long long h[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h, c, 16, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *nodes_out = h[0];
  *prims_out = h[1];
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_render_configure(TutuCtx* ctx, uint64_t paths_in_flight, int lanes, int profile_stages) {
  if (!ctx) return fail(nullptr, TUTU_E_INVALID, "null context");
  if (paths_in_flight > ((uint64_t)1 << 31)) return fail(ctx, TUTU_E_INVALID, "paths_in_flight too large");
  if (lanes < 0 || lanes > 8) return fail(ctx, TUTU_E_INVALID, "lanes must be 0 (default) .. 8");
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->paths_in_flight_cfg = paths_in_flight;
  ctx->lanes_cfg = lanes;
  ctx->profile_stages = profile_stages;
  return TUTU_OK;
}

extern "C" int tutu_render_pipeline(TutuCtx* ctx, int pipeline) {
  if (!ctx) return fail(nullptr, TUTU_E_INVALID, "null context");
  if (pipeline < 0 || pipeline > 2) return fail(ctx, TUTU_E_INVALID, "pipeline must be 0 (automatic), 1 (wavefront) or 2 (register-resident)");
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->pipeline_cfg = pipeline;
  return TUTU_OK;
}

extern "C" int tutu_render_path_accumulate_device(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count,
                                                  uint64_t seed, float* d_accum, void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!d_accum) return fail(ctx, TUTU_E_INVALID, "tutu_render_path_accumulate_device: null accumulation buffer");
  if (int rc = check_pipeline(ctx)) return rc;
  wf_render(ctx, sample_begin, sample_count, seed, d_accum, stream ? (cudaStream_t)stream : ctx->stream);
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_finalize_device(TutuCtx* ctx, const float* d_accum, float inv_spp, float* d_rgb_out,
                                    void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!d_accum || !d_rgb_out) return fail(ctx, TUTU_E_INVALID, "tutu_finalize_device: null buffer");
  const size_t n = (size_t)ctx->flat.raygen.width * ctx->flat.raygen.height * 3;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  wf_finalize<<<ctx->sm_count * 4, 256, 0, s>>>(d_accum, inv_spp, d_rgb_out, n);
  CUDA_TRY(cudaGetLastError());
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_render_path(TutuCtx* ctx, uint32_t spp, uint64_t seed, float* rgb_out) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!rgb_out || spp == 0) return fail(ctx, TUTU_E_INVALID, "tutu_render_path: null output or spp == 0");
  if (int rc = check_pipeline(ctx)) return rc;
  const size_t n = (size_t)ctx->flat.raygen.width * ctx->flat.raygen.height * 3;
  cudaStream_t s = ctx->stream;
  ctx->d_accum.ensure(n * sizeof(float));
  ctx->d_rgb.ensure(n * sizeof(float));
  CUDA_TRY(cudaMemsetAsync(ctx->d_accum.p, 0, n * sizeof(float), s));
  wf_render(ctx, 0, spp, seed, ctx->d_accum.as<float>(), s);
  wf_finalize<<<ctx->sm_count * 4, 256, 0, s>>>(ctx->d_accum.as<float>(), 1.f / (float)spp, ctx->d_rgb.as<float>(), n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(rgb_out, ctx->d_rgb.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_render_bdpt_accumulate_device(TutuCtx* ctx, uint32_t sample_begin, uint32_t sample_count,
                                                  uint64_t seed, float* d_accum, void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!d_accum) return fail(ctx, TUTU_E_INVALID, "tutu_render_bdpt_accumulate_device: null accumulation buffer");
  bdpt_render(ctx, sample_begin, sample_count, seed, d_accum, stream ? (cudaStream_t)stream : ctx->stream);
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_finalize_bdpt_device(TutuCtx* ctx, const float* d_accum, float inv_spp, float* d_rgb_out,
                                         void* stream) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!d_accum || !d_rgb_out) return fail(ctx, TUTU_E_INVALID, "tutu_finalize_bdpt_device: null buffer");
  const size_t npix = (size_t)ctx->flat.raygen.width * ctx->flat.raygen.height;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  const float* bk = ctx->flat.bkgcolor;
  bdpt_finalize<<<ctx->sm_count * 4, 256, 0, s>>>(d_accum, inv_spp, bk[0], bk[1], bk[2], d_rgb_out, npix);
  CUDA_TRY(cudaGetLastError());
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_render_bdpt(TutuCtx* ctx, uint32_t spp, uint64_t seed, float* rgb_out) {
  API_BEGIN(ctx)
  if (int rc = check_scene(ctx)) return rc;
  if (!rgb_out || spp == 0) return fail(ctx, TUTU_E_INVALID, "tutu_render_bdpt: null output or spp == 0");
  const size_t npix = (size_t)ctx->flat.raygen.width * ctx->flat.raygen.height;
  const size_t n = npix * 3;
  cudaStream_t s = ctx->stream;
  ctx->d_accum.ensure(n * sizeof(float));
  ctx->d_rgb.ensure(n * sizeof(float));
  CUDA_TRY(cudaMemsetAsync(ctx->d_accum.p, 0, n * sizeof(float), s));
  bdpt_render(ctx, 0, spp, seed, ctx->d_accum.as<float>(), s);
  const float* bk = ctx->flat.bkgcolor;
  bdpt_finalize<<<ctx->sm_count * 4, 256, 0, s>>>(ctx->d_accum.as<float>(), 1.f / (float)spp, bk[0], bk[1], bk[2],
                                                  ctx->d_rgb.as<float>(), npix);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(rgb_out, ctx->d_rgb.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_quantize_device(TutuCtx* ctx, const float* d_rgb, uint64_t n_pixels, float gamma, uint8_t* d_out,
                                    void* stream) {
  API_BEGIN(ctx)
  if (n_pixels == 0) return TUTU_OK;
  if (!d_rgb || !d_out) return fail(ctx, TUTU_E_INVALID, "tutu_quantize_device: null buffer");
  if (((uintptr_t)d_rgb & 15u) || ((uintptr_t)d_out & 3u))
    return fail(ctx, TUTU_E_INVALID, "tutu_quantize_device: rgb must be 16-byte and out 4-byte aligned");
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  k_quantize<<<ctx->sm_count * 8, 256, 0, s>>>(d_rgb, (size_t)n_pixels * 3, gamma, d_out);
  CUDA_TRY(cudaGetLastError());
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_quantize(TutuCtx* ctx, const float* rgb, uint64_t n_pixels, float gamma, uint8_t* out) {
  API_BEGIN(ctx)
  if (n_pixels == 0) return TUTU_OK;
  if (!rgb || !out) return fail(ctx, TUTU_E_INVALID, "tutu_quantize: null buffer");
  const size_t n = (size_t)n_pixels * 3;
  cudaStream_t s = ctx->stream;
  ctx->d_rgb.ensure(n * sizeof(float));
  ctx->d_blocked.ensure(n);
  CUDA_TRY(cudaMemcpyAsync(ctx->d_rgb.p, rgb, n * sizeof(float), cudaMemcpyHostToDevice, s));
  k_quantize<<<ctx->sm_count * 8, 256, 0, s>>>(ctx->d_rgb.as<float>(), n, gamma, ctx->d_blocked.as<uint8_t>());
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, ctx->d_blocked.p, n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TUTU_OK;
  API_END(ctx)
}

// ---- Postprocessor (postprocess.cuh) ---------------------------------------------------------------
namespace {
// Enqueues `mode` on `s`: src -> dst (both device, width*height*3 floats, not aliased); scratch buffers are
// the ctx's.  The kernel sequence is the statement sequence of Postprocessor::performPostProcess (:37-58).
void post_enqueue(TutuCtx* ctx, const float* src, uint32_t width, uint32_t height, int mode, const TutuPostParams& p,
                  float* dst, cudaStream_t s) {
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  const int w = (int)width, h = (int)height;
  const int grid = ctx->sm_count * 8;
  PostWeights pw{};
  if (mode == TUTU_POST_BLUR || mode == TUTU_POST_BLOOM || mode == TUTU_POST_HDR_BLOOM) {
    pw.n = p.kernel_size;
    post_gaussian_weights(p.kernel_size, p.stddev, pw.g, &pw.sum, &pw.start);
  }
  auto blur = [&](const float* in, float* tmp, float* out) {  // vertical pass, then horizontal (src = res between)
    pp_blur<true><<<grid, 256, 0, s>>>(in, w, h, pw, tmp);
    pp_blur<false><<<grid, 256, 0, s>>>(tmp, w, h, pw, out);
  };
  switch (mode) {
    case TUTU_POST_EXTRACT:
      pp_extract<<<grid, 256, 0, s>>>(src, w, h, p.emissive_norm, p.strength, dst);
      break;
    case TUTU_POST_BLUR:
      ctx->d_post[0].ensure(bytes);
      blur(src, ctx->d_post[0].as<float>(), dst);
      break;
    case TUTU_POST_HDR:
      pp_combine<true><<<grid, 256, 0, s>>>(src, nullptr, w, h, p.exposure, dst);
      break;
    case TUTU_POST_BLOOM:
    case TUTU_POST_HDR_BLOOM: {
      for (int k = 0; k < 3; ++k) ctx->d_post[k].ensure(bytes);
      float* a = ctx->d_post[0].as<float>();
      float* b = ctx->d_post[1].as<float>();
      float* tmp = ctx->d_post[2].as<float>();
      pp_extract<<<grid, 256, 0, s>>>(src, w, h, p.emissive_norm, p.strength, a);
      for (int k = 0; k < 1 + p.gaussian_loops; ++k) {
        blur(a, tmp, b);
        std::swap(a, b);
      }
      if (mode == TUTU_POST_BLOOM)
        pp_combine<false><<<grid, 256, 0, s>>>(src, a, w, h, p.exposure, dst);
      else
        pp_combine<true><<<grid, 256, 0, s>>>(src, a, w, h, p.exposure, dst);
    } break;
  }
  CUDA_TRY(cudaGetLastError());
}

int post_check(TutuCtx* ctx, const void* in, const void* out, uint32_t width, uint32_t height, int mode, const TutuPostParams* params,
               TutuPostParams* p) {
  if (!in || !out) return fail(ctx, TUTU_E_INVALID, "tutu_postprocess: null buffer");
  if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 30))
    return fail(ctx, TUTU_E_INVALID, "tutu_postprocess: bad image size");
  if (mode < TUTU_POST_EXTRACT || mode > TUTU_POST_HDR_BLOOM) return fail(ctx, TUTU_E_INVALID, "tutu_postprocess: unknown mode");
  if (params)
    *p = *params;
  else
    tutu_post_params_default(p);
  if (p->kernel_size < 1 || p->kernel_size > kPostMaxKernel || p->gaussian_loops < 0 || p->gaussian_loops > 64 || !(p->stddev > 0.f))
    return fail(ctx, TUTU_E_INVALID, "tutu_postprocess: kernel_size must be 1..64, gaussian_loops 0..64, stddev > 0");
  return TUTU_OK;
}
}  // namespace

extern "C" int tutu_postprocess_device(TutuCtx* ctx, const float* d_rgb, uint32_t width, uint32_t height, int mode,
                                       const TutuPostParams* params, float* d_rgb_out, void* stream) {
  API_BEGIN(ctx)
  TutuPostParams p;
  if (int rc = post_check(ctx, d_rgb, d_rgb_out, width, height, mode, params, &p)) return rc;
  if (d_rgb == d_rgb_out) return fail(ctx, TUTU_E_INVALID, "tutu_postprocess_device: output must not alias the input");
  post_enqueue(ctx, d_rgb, width, height, mode, p, d_rgb_out, stream ? (cudaStream_t)stream : ctx->stream);
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_postprocess(TutuCtx* ctx, const float* rgb, uint32_t width, uint32_t height, int mode,
                                const TutuPostParams* params, float* rgb_out) {
  API_BEGIN(ctx)
  TutuPostParams p;
  if (int rc = post_check(ctx, rgb, rgb_out, width, height, mode, params, &p)) return rc;
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  cudaStream_t s = ctx->stream;
  ctx->d_post[3].ensure(bytes);
  ctx->d_post[4].ensure(bytes);
  CUDA_TRY(cudaMemcpyAsync(ctx->d_post[3].p, rgb, bytes, cudaMemcpyHostToDevice, s));
  post_enqueue(ctx, ctx->d_post[3].as<float>(), width, height, mode, p, ctx->d_post[4].as<float>(), s);
  CUDA_TRY(cudaMemcpyAsync(rgb_out, ctx->d_post[4].p, bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TUTU_OK;
  API_END(ctx)
}

extern "C" int tutu_render_stats(const TutuCtx* ctx, TutuRenderStats* out) {
  if (!ctx || !out) return fail(nullptr, TUTU_E_INVALID, "tutu_render_stats: null argument");
  *out = ctx->stats;
  return TUTU_OK;
}
