// bdpt.cuh — the BDPT integrator (reference include/BDPT.hpp) as a wavefront pipeline.
//
//   bdpt_start     camera vertex + primary ray, light vertex 0 (sampleLight + sampleLightDir) and
//                  its first ray                                        BDPT.hpp:707-741, 296-330
//   q_extend       closest hit for the walk queue (eye and light sub-paths share one queue)
//   bdpt_vertex    one step of buildEyePath / buildLightPath: texture, sampleDirection, fwd/rev
//                  pdf, G, store the compact vertex, next ray           BDPT.hpp:236-292, 334-389
//   bdpt_connect   every (s,t) strategy of one path length: unweighted contribution, MISweight,
//                  any-hit ray into the shadow queue (s = 0 adds directly)  BDPT.hpp:752-886, 70-222
//   q_shadow_add   isShadowRayBlocked for the queue + atomic add / t = 1 splat (BDPT.hpp:819-824)
//   bdpt_finalize  bkgcolor + sum * SPP_inv (Camera.hpp:28, BDPT.hpp:891)
//
// A batch of B samples is walked for the fixed 7 iterations (no host polling: iteration `it`
// handles the vertices with index it+1 of both sub-path kinds), then the seven path lengths are
// connected one launch each.  Vertices live in HBM as float4 SoA [slot][sample]: eye vertices
// 1..7 (the camera vertex 0 is implicit) and light vertices 0..6.
//
// Reference quirks kept: t = 1 uses Ng for the offset side (BDPT.hpp:803, the shipped
// MULTITHREAD == 1 worker); (s=1,t=1) never contributes (the light vertex is emissive, :790);
// MISweight returns 0 below MIN_DIVISOR (:218); only contrib.x is NaN-tested (:776,810,879); a
// missed primary ray ends the pixel (:733-734); the UNLIT test at :767 is unreachable (an UNLIT
// first hit fails sampleDirection, so epverts.size() < 2 at :750).
#pragma once
#include "tutu_internal.hpp"
#include "wavefront.cuh"

namespace tutu {

constexpr int kBdptMaxLen = 7;      // MAX_PATHLENGTH, BDPT.hpp:8
constexpr int kBdptEyeSlots = 7;    // stored eye vertices 1..7
constexpr int kBdptLightSlots = 7;  // stored light vertices 0..6
constexpr int kBdptSlots = kBdptEyeSlots + kBdptLightSlots;
constexpr uint32_t kRngEye = 32u, kRngLight = 64u;  // Philox counter "depth" of the two walks
constexpr uint32_t kKindEye = 0u, kKindLight = 1u;
constexpr uint32_t kDeadEntry = 0xFFFFFFFFu;

using BdptCam = BdptCamConsts;  // computed on the host (host_scene.cpp: compute_bdpt_cam)

struct BdptCtl {
  unsigned n_cur, n_next, n_shadow, pad;
  unsigned long long cursor_extend, cursor_shadow;
  unsigned long long sum_extend, sum_shadow, connections;
};

struct BdptBuffers {
  // vertex store, index = slot * cap + sample; slots [0,7) eye vertices 1..7, [7,14) light 0..6
  float4* vP;    // pos.xyz, bits(leaf slot code)
  float4* vNg;   // Ng.xyz, -
  float4* vNs;   // Ns.xyz (after normal mapping), -
  float4* vT;    // throughput.xyz, -
  float4* vM;    // diffuse.xyz, roughness (after textureModify)
  float4* vMis;  // fwdPdf, revPdf, G, bits(material id | isDelta << 31)
  float* vMet;   // metallic (after textureModify)
  unsigned char* nE;  // epverts.size() per sample (camera included), 0 = primary ray missed
  unsigned char* nL;  // lpverts.size() per sample
  // walk queue (ping-pong)
  float4* q_o[2];   // o.xyz, bits(sample slot)
  float4* q_d[2];   // d.xyz, bits(kind << 8 | vertex index), kDeadEntry = skip
  float4* q_tp[2];  // throughput carried to the next vertex .xyz, -
  float4* hit;
  // shadow queue
  float4* sh_o;  // o.xyz, dist
  float4* sh_d;  // d.xyz, bits(pixel)
  float4* sh_c;  // weighted contribution .xyz, -
  BdptCtl* ctl;
  float* accum;
  unsigned cap;
};

// ---- camera helpers ----------------------------------------------------------------------------
__device__ __forceinline__ f3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }

__device__ __forceinline__ int worldPos2PixelIndex(const BdptCam& c, f3 p) {  // Camera.hpp:51-78
  // Exact fp32 in the reference's order (Mat4f * Vector4f, Vector.hpp:289-296; w = 1): the last
  // pixel row / column projects to raster coordinate H / W up to rounding (pixel centres are spaced
  // (ll-ul)/(H-1) but offset by (ll-ul)/(2H), BDPT.hpp:416-418), so whether We() of those pixels is
  // zero is decided by the last bit here.
  const float* m = c.w2r;
  auto row = [&](int r) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, m[4 * r]), __fmul_rn(p.y, m[4 * r + 1])), __fmul_rn(p.z, m[4 * r + 2])),
                     __fmul_rn(1.f, m[4 * r + 3]));
  };
  const float rx = row(0), ry = row(1), rw = row(3);
  const int x = (int)__fsub_rn(__fdiv_rn(rx, rw), 0.5f);
  const int y = (int)__fsub_rn(__fdiv_rn(ry, rw), 0.5f);
  if (x < 0 || x >= c.width || y < 0 || y >= c.height) return -1;
  return x + c.width * y;
}

// We(), IIntegrator.hpp:233-248; *index_out = the pixel the point projects to
__device__ __forceinline__ float We(const BdptCam& c, f3 pos, int* index_out) {
  const f3 inter2cam = normalized(ld3(c.eye) - pos);
  const int index = worldPos2PixelIndex(c, pos);
  *index_out = index;
  if (index < 0 || index >= c.width * c.height) return 0.f;
  const float cosCamera = fabsf(dot(ld3(c.fwd), -inter2cam));
  const float distPixel2Cam = c.imagePlaneDist / cosCamera;
  return distPixel2Cam * distPixel2Cam * c.lensAreaInv * c.filmPlaneAreaInv / (cosCamera * cosCamera);
}

__device__ __forceinline__ float Geo(f3 p1, f3 n1, f3 p2, f3 n2) {  // IIntegrator.hpp:223-230
  f3 d = p2 - p1;
  const float dis2 = norm2(d);
  d = normalized(d);
  return fabsf(dot(d, n1)) * fabsf(dot(-d, n2)) / dis2;
}

// pixel centre and primary direction, BDPT.hpp:695-702, exact fp32 in the reference's order
__device__ __forceinline__ void bdpt_pixel(const BdptCam& c, uint32_t pixel, f3& pixelPos, f3& rayDir) {
  const float x = (float)(pixel % (uint32_t)c.width), y = (float)(pixel / (uint32_t)c.width);
  float p[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    p[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(c.ul[k], __fmul_rn(x, c.dh[k])), __fmul_rn(y, c.dv[k])), c.coh[k]), c.cov[k]);
  pixelPos = mk(p[0], p[1], p[2]);
  float dx = __fsub_rn(p[0], c.eye[0]), dy = __fsub_rn(p[1], c.eye[1]), dz = __fsub_rn(p[2], c.eye[2]);
  normalize_rn(dx, dy, dz);
  rayDir = mk(dx, dy, dz);
}

// camera vertex pdfs, BDPT.hpp:721-730
__device__ __forceinline__ void bdpt_cam_pdfs(const BdptCam& c, f3 pixelPos, f3 rayDir, float& fwdPdf, float& tp0) {
  const float wi_n_cos = fabsf(dot(rayDir, ld3(c.fwd)));
  const float d2 = norm2(pixelPos - ld3(c.eye));
  fwdPdf = d2 * c.filmPlaneAreaInv / wi_n_cos;
  fwdPdf = fwdPdf / wi_n_cos;
  const float pdfCam_w = d2 * c.lensAreaInv * c.filmPlaneAreaInv / wi_n_cos;
  tp0 = 1.f * wi_n_cos / pdfCam_w;
}

// sampleLightDir, IIntegrator.hpp:195-220
__device__ __forceinline__ bool sampleLightDir(f3 N, float r1, float r2, float& dirPdf, f3& out) {
  const float cosTheta = sqrtf(r1);
  const float phi = 2 * T_PI * r2;
  const float sinTheta = sqrtf(fmaxf(0.f, 1 - r1));
  float sp, cp;
  sincosf(phi, &sp, &cp);
  const f3 dir = normalized(mk(cp * sinTheta, sp * sinTheta, cosTheta));
  const f3 res = SphereLocal2world(N, dir);
  if (dot(normalized(res), N) < 0) return false;
  dirPdf = 0.f;
  if (dot(res, N) > 0.0f) dirPdf = dot(res, N) / T_PI;
  out = res;
  return true;
}

// ---- control -----------------------------------------------------------------------------------
__global__ void bdpt_ctl_begin(BdptCtl* ctl, unsigned n_entries) {
  ctl->n_cur = n_entries;
  ctl->n_next = 0;
  ctl->n_shadow = 0;
  ctl->cursor_extend = 0;
  ctl->cursor_shadow = 0;
  ctl->sum_extend += n_entries;
}
__global__ void bdpt_ctl_after_walk(BdptCtl* ctl) {
  ctl->n_cur = ctl->n_next;
  ctl->n_next = 0;
  ctl->cursor_extend = 0;
  ctl->sum_extend += ctl->n_cur;
}
__global__ void bdpt_ctl_after_shadow(BdptCtl* ctl) {
  ctl->sum_shadow += ctl->n_shadow;
  ctl->n_shadow = 0;
  ctl->cursor_shadow = 0;
}

// ---- generic queue tracers (same packet scheme as wf_extend / wf_shadow) --------------------------
// KIND: 0 = binary trees, traversal stack in shared memory, 1 = small scene (flat tests), 2 = binary trees, stack in local
// memory (as in wavefront.cuh); the wide-tree tracers are in trace_kernels.cu
template <int KIND>
__global__ void __launch_bounds__(256)
q_extend(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, const float4* __restrict__ ro,
         const float4* __restrict__ rd, float4* __restrict__ hit, const unsigned* __restrict__ n_ptr,
         unsigned long long* cursor) {
  extern __shared__ unsigned long long s_stack[];  // KIND 0: traversal stack (trace.cuh: SharedStack)
  const unsigned n = *n_ptr;
  const unsigned lane = threadIdx.x & 31u;
  if constexpr (KIND == 0 || KIND == 2) {
    if (sc.queue_lanes) {  // persistent lanes, phase-separated steps (trace.cuh: trace_queue_lanes)
      auto store = [&](unsigned i, const Hit& h, bool) { __stcs(hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot))); };
      auto load = [&](unsigned i, Ray& r, float& dis) {
        const float4 o = __ldcs(ro + i), d = __ldcs(rd + i);
        if (__float_as_uint(d.w) == kDeadEntry) {
          __stcs(hit + i, make_float4(FLT_MAX, 0.f, 0.f, __int_as_float(-1)));
          return false;
        }
        r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
        dis = 0.f;
        return true;
      };
      if constexpr (KIND == 0) {
        SharedStack<false> st;
        st.base = s_stack + threadIdx.x;
        st.stride = blockDim.x;
        trace_queue_lanes<false>(sc, n, cursor, st, load, store);
      } else {
        LocalStack<false> st;
        trace_queue_lanes<false>(sc, n, cursor, st, load, store);
      }
      return;
    }
  }
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned i = (unsigned)base + k + lane;
      if (i < n) {
        const float4 o = __ldcs(ro + i);
        const float4 d = __ldcs(rd + i);
        Hit h;
        h.t = FLT_MAX, h.u = 0.f, h.v = 0.f, h.slot = -1;
        if (__float_as_uint(d.w) != kDeadEntry) {
          if constexpr (KIND == 1)
            traverse_small<false>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h);
          else if constexpr (KIND == 2)
            traverse_structured<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h);
          else
            traverse_shared<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h, s_stack);
        }
        __stcs(hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
      }
      __syncwarp();
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(256)
q_shadow_add(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, const float4* __restrict__ so,
             const float4* __restrict__ sd, const float4* __restrict__ scn, float* __restrict__ accum,
             const unsigned* __restrict__ n_ptr, unsigned long long* cursor) {
  extern __shared__ unsigned long long s_stack[];  // KIND 0: traversal stack, 32-bit entries (trace.cuh: SharedStack<true>)
  const unsigned n = *n_ptr;
  const unsigned lane = threadIdx.x & 31u;
  if constexpr (KIND == 0 || KIND == 2) {
    if (sc.queue_lanes) {  // persistent lanes, phase-separated steps (trace.cuh: trace_queue_lanes)
      auto load = [&](unsigned j, Ray& r, float& dis) {
        const float4 o = __ldcs(so + j), d = __ldcs(sd + j);
        r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
        dis = o.w;
        return true;
      };
      auto store = [&](unsigned j, const Hit&, bool blocked) {
        if (!blocked) {
          const float4 c = __ldcs(scn + j);
          float* p = accum + (size_t)__float_as_uint(__ldcs(sd + j).w) * 3;
          atomicAdd(p + 0, c.x);
          atomicAdd(p + 1, c.y);
          atomicAdd(p + 2, c.z);
        }
      };
      if constexpr (KIND == 0) {
        SharedStack<true> st;
        st.base = reinterpret_cast<unsigned*>(s_stack) + threadIdx.x;
        st.stride = blockDim.x;
        trace_queue_lanes<true>(sc, n, cursor, st, load, store);
      } else {
        LocalStack<true> st;
        trace_queue_lanes<true>(sc, n, cursor, st, load, store);
      }
      return;
    }
  }
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kPacketRays);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base >= n) return;
#pragma unroll 1
    for (unsigned k = 0; k < kPacketRays; k += 32u) {
      const unsigned j = (unsigned)base + k + lane;
      if (j < n) {
        const float4 o = __ldcs(so + j);
        const float4 d = __ldcs(sd + j);
        Hit h;
        bool blocked;
        if constexpr (KIND == 1)
          blocked = traverse_small<true>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h);
        else if constexpr (KIND == 2)
          blocked = traverse_structured<true>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h);
        else
          blocked = traverse_shared<true>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h, s_stack);
        if (!blocked) {
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

// ---- vertices ----------------------------------------------------------------------------------
struct Vtx {
  f3 pos, Ng, Ns, tp;
  Mat m;
  uint32_t code;  // leaf slot code (slot | sphere bit)
};

__device__ __forceinline__ void store_vtx(const BdptBuffers& b, int slot, unsigned i, f3 pos, uint32_t code, f3 Ng, f3 Ns,
                                          f3 tp, const Mat& m, int mat_id, float fwd, float rev, float G, bool delta) {
  const size_t at = (size_t)slot * b.cap + i;
  b.vP[at] = make_float4(pos.x, pos.y, pos.z, __uint_as_float(code));
  b.vNg[at] = make_float4(Ng.x, Ng.y, Ng.z, 0.f);
  b.vNs[at] = make_float4(Ns.x, Ns.y, Ns.z, 0.f);
  b.vT[at] = make_float4(tp.x, tp.y, tp.z, 0.f);
  b.vM[at] = make_float4(m.diffuse.x, m.diffuse.y, m.diffuse.z, m.roughness);
  b.vMis[at] = make_float4(fwd, rev, G, __uint_as_float((uint32_t)mat_id | (delta ? 0x80000000u : 0u)));
  b.vMet[at] = m.metallic;
}

__device__ __forceinline__ Vtx load_vtx(const DevScene& sc, const BdptBuffers& b, int slot, unsigned i) {
  const size_t at = (size_t)slot * b.cap + i;
  Vtx v;
  const float4 p = b.vP[at], ng = b.vNg[at], ns = b.vNs[at], t = b.vT[at], mm = b.vM[at], mis = b.vMis[at];
  v.pos = mk(p.x, p.y, p.z);
  v.code = __float_as_uint(p.w);
  v.Ng = mk(ng.x, ng.y, ng.z);
  v.Ns = mk(ns.x, ns.y, ns.z);
  v.tp = mk(t.x, t.y, t.z);
  v.m = load_material(sc, (int)(__float_as_uint(mis.w) & 0x7FFFFFFFu));
  v.m.diffuse = mk(mm.x, mm.y, mm.z);
  v.m.roughness = mm.w;
  v.m.metallic = b.vMet[at];
  return v;
}

// epverts[ti] / lpverts[li] accessors (ti = 0 is the camera vertex)
__device__ __forceinline__ f3 eye_pos(const BdptCam& c, const BdptBuffers& b, unsigned i, int ti) {
  if (ti == 0) return ld3(c.eye);
  const float4 p = b.vP[(size_t)(ti - 1) * b.cap + i];
  return mk(p.x, p.y, p.z);
}
__device__ __forceinline__ f3 light_pos(const BdptBuffers& b, unsigned i, int li) {
  const float4 p = b.vP[(size_t)(kBdptEyeSlots + li) * b.cap + i];
  return mk(p.x, p.y, p.z);
}
struct MisRec {
  float fwd, rev, G;
  bool delta;
};
__device__ __forceinline__ MisRec eye_mis(const BdptCam& c, const BdptBuffers& b, unsigned i, int ti, float camFwd) {
  if (ti == 0) return MisRec{camFwd, c.lensAreaInv, 0.f, false};
  const float4 m = b.vMis[(size_t)(ti - 1) * b.cap + i];
  return MisRec{m.x, m.y, m.z, (__float_as_uint(m.w) & 0x80000000u) != 0u};
}
__device__ __forceinline__ MisRec light_mis(const BdptBuffers& b, unsigned i, int li) {
  const float4 m = b.vMis[(size_t)(kBdptEyeSlots + li) * b.cap + i];
  return MisRec{m.x, m.y, m.z, (__float_as_uint(m.w) & 0x80000000u) != 0u};
}

// ---- start -------------------------------------------------------------------------------------
// entry i = eye walk of sample i (neighbouring pixels share a warp), entry n + i = its light walk
__global__ void __launch_bounds__(256)
bdpt_start(const __grid_constant__ DevScene sc, const __grid_constant__ BdptCam cam, BdptBuffers b,
           unsigned long long first_path, unsigned n, unsigned sample_begin, uint64_t seed) {
  const unsigned npix = (unsigned)cam.width * (unsigned)cam.height;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long g = first_path + i;
    const uint32_t pixel = (uint32_t)(g % npix);
    const uint32_t sample = sample_begin + (uint32_t)(g / npix);
    f3 pixelPos, rayDir;
    bdpt_pixel(cam, pixel, pixelPos, rayDir);
    float camFwd, tp0;
    bdpt_cam_pdfs(cam, pixelPos, rayDir, camFwd, tp0);
    b.q_o[0][i] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], __uint_as_float(i));
    b.q_d[0][i] = make_float4(rayDir.x, rayDir.y, rayDir.z, __uint_as_float((kKindEye << 8) | 1u));
    b.q_tp[0][i] = make_float4(tp0, tp0, tp0, 0.f);
    b.nE[i] = 0;  // becomes >= 1 once the primary ray hits (bdpt_vertex)
    b.nL[i] = 0;
    // ---- buildLightPath up to the first ray, BDPT.hpp:296-326 ----
    float4 lo = make_float4(0, 0, 0, __uint_as_float(i)), ld = make_float4(0, 0, 0, __uint_as_float(kDeadEntry));
    float4 ltp = make_float4(0, 0, 0, 0);
    if (sc.n_lights > 0) {
      const Rand4 r0 = draw4(seed, pixel, sample, kRngLight, 0u);
      const Rand4 r1 = draw4(seed, pixel, sample, kRngLight, 1u);
      const int size = sc.n_lights;
      int index = (int)(r0.u[0] * (size - 1) + 0.4999f);  // IIntegrator.hpp:184
      if (size == 1) index = 0;
      const LightSample ls = sample_light(sc, r0.u[0], r0.u[1], r0.u[2]);
      const float4 l1 = __ldg(sc.lights + 8 * (size_t)index + 1), l2 = __ldg(sc.lights + 8 * (size_t)index + 2);
      const float4 l4 = __ldg(sc.lights + 8 * (size_t)index + 4);
      const bool sphere = __float_as_int(l1.w) == TUTU_PRIM_SPHERE;
      const uint32_t code = (uint32_t)__float_as_int(l2.w) | (sphere ? kSphereBit : 0u);
      const int mat_id = __float_as_int(l4.w);
      const float pickpdf = ls.pdf;
      float dirPdf;
      f3 wi;
      if (sampleLightDir(ls.Ns, r0.u[3], r1.u[0], dirPdf, wi)) {  // samplePoint sets Ng = Ns (Triangle.hpp:126-128)
        wi = normalized(wi);
        const float wi_n_cos = fabsf(dot(wi, ls.Ns));
        const Mat m = load_material(sc, mat_id);
        const f3 tp = mk(fdiv(1.f, pickpdf));
        store_vtx(b, kBdptEyeSlots + 0, i, ls.pos, code, ls.Ns, ls.Ns, tp, m, mat_id, dirPdf / wi_n_cos, pickpdf, 0.f, false);
        b.nL[i] = 1;
        const f3 tp2 = tp * (wi_n_cos / dirPdf);
        const f3 orig = ls.pos + ls.Ns * T_EPSILON;
        lo = make_float4(orig.x, orig.y, orig.z, __uint_as_float(i));
        ld = make_float4(wi.x, wi.y, wi.z, __uint_as_float((kKindLight << 8) | 1u));
        ltp = make_float4(tp2.x, tp2.y, tp2.z, 0.f);
      }
    }
    b.q_o[0][n + i] = lo;
    b.q_d[0][n + i] = ld;
    b.q_tp[0][n + i] = ltp;
  }
}

// ---- one walk step -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
bdpt_vertex(const __grid_constant__ DevScene sc, const __grid_constant__ BdptCam cam, BdptBuffers b, int cur,
            unsigned long long first_path, unsigned sample_begin, uint64_t seed) {
  const int nxt = cur ^ 1;
  const unsigned n = b.ctl->n_cur;
  const unsigned npix = (unsigned)cam.width * (unsigned)cam.height;
  for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned j = base + threadIdx.x;
    bool cont = false;
    float4 no = make_float4(0, 0, 0, 0), nd = no, ntp = no;
    if (j < n) {
      const float4 o = __ldcs(b.q_o[cur] + j);
      const float4 d = __ldcs(b.q_d[cur] + j);
      const float4 hit = __ldcs(b.hit + j);
      const uint32_t tag = __float_as_uint(d.w);
      if (tag != kDeadEntry && __float_as_int(hit.w) >= 0) {
        const unsigned i = __float_as_uint(o.w);
        const uint32_t kind = tag >> 8, k = tag & 0xFFu;  // this hit becomes vertex k of its sub-path
        const float4 tp4 = __ldcs(b.q_tp[cur] + j);
        f3 tp = mk(tp4.x, tp4.y, tp4.z);
        const Ray ray{o.x, o.y, o.z, d.x, d.y, d.z};
        Surf s = load_surface(sc, ray, hit);
        const bool emissive = s.m.emission.x || s.m.emission.y || s.m.emission.z;
        // the first light-walk hit on an emitter ends the walk before a vertex exists (BDPT.hpp:329-330)
        if (!(kind == kKindLight && k == 1u && emissive)) {
          if (kind == kKindEye && k == 1u) b.nE[i] = 1;  // primary ray hit: the sample is alive (:733)
          if (s.textured) {
            const TexMod tm = texture_modify(sc, s.slot, s.sphere, s.tu, s.tv, s.Ng,
                                             TexMod{s.m.diffuse, s.Ns, s.m.roughness, s.m.metallic});
            s.m.diffuse = tm.diffuse, s.Ns = tm.Ns, s.m.roughness = tm.roughness, s.m.metallic = tm.metallic;
          }
          const unsigned long long g = first_path + i;
          const uint32_t pixel = (uint32_t)(g % npix);
          const uint32_t sample = sample_begin + (uint32_t)(g / npix);
          const Rand4 rn = draw4(seed, pixel, sample, (kind == kKindEye ? kRngEye : kRngLight) + k, 0u);
          f3 wi = mk(ray.dx, ray.dy, ray.dz);
          const f3 wo = -wi;
          const int ok = sampleDirection(s.m, wo, s.Ns, wi, sc.eta, rn.u[0], rn.u[1], rn.u[2]);
          if (ok & 1) {
            const bool TIR = (ok & 2) != 0;
            wi = normalized(wi);
            float dirPdf = mat_pdf_eval(s.m, wi, wo, s.Ns, sc.eta, s.m.eta);
            if (TIR) {
              wi = normalized(getReflectionDir(wo, s.Ns));
              dirPdf = 1;
            }
            if (dirPdf != 0) {
              const float cosv = fabsf(dot(wi, s.Ng));
              const float fwd = dirPdf / cosv;
              float rev;
              bool delta;
              if (s.m.type == TUTU_MAT_PERFECT_REFLECTIVE || s.m.type == TUTU_MAT_PERFECT_REFRACTIVE) {
                rev = fwd;
                delta = true;
              } else {
                rev = mat_pdf_eval(s.m, wo, wi, s.Ns, sc.eta, s.m.eta) / fabsf(dot(wo, s.Ng));
                delta = false;
              }
              // G with the previous vertex of the same sub-path
              f3 ppos, pNg;
              if (kind == kKindEye && k == 1u) {
                ppos = ld3(cam.eye), pNg = ld3(cam.fwd);
              } else {
                const int pslot = kind == kKindEye ? (int)k - 2 : kBdptEyeSlots + (int)k - 1;
                const float4 pp = b.vP[(size_t)pslot * b.cap + i], pn = b.vNg[(size_t)pslot * b.cap + i];
                ppos = mk(pp.x, pp.y, pp.z), pNg = mk(pn.x, pn.y, pn.z);
              }
              const float G = Geo(ppos, pNg, s.pos, s.Ng);
              const uint32_t code = s.slot | (s.sphere ? kSphereBit : 0u);
              const int mat_id = (int)(__float_as_uint(__ldg(sc.shade + 4 * (size_t)s.slot + 3).w) & 0x3FFFFFFFu);
              const int slot = kind == kKindEye ? (int)k - 1 : kBdptEyeSlots + (int)k;
              store_vtx(b, slot, i, s.pos, code, s.Ng, s.Ns, tp, s.m, mat_id, fwd, rev, G, delta);
              (kind == kKindEye ? b.nE : b.nL)[i] = (unsigned char)(k + 1u);
              // the walk goes on unless the vertex is emissive, its pdf is below MIN_DIVISOR, or the
              // sub-path is full (eye: 8 vertices, light: 7; BDPT.hpp:236,334)
              const uint32_t limit = kind == kKindEye ? (uint32_t)kBdptMaxLen : (uint32_t)kBdptMaxLen - 1u;
              if (!emissive && !(dirPdf < T_MIN_DIVISOR) && k < limit) {
                const f3 bsdf = kind == kKindEye ? BxDF(s.m, wi, wo, s.Ng, s.Ns, sc.eta, TIR)
                                                 : BxDF_adjoint(s.m, wi, wo, s.Ng, s.Ns, sc.eta, TIR);
                tp = tp * bsdf * (cosv / dirPdf);
                const bool rayInside = dot(s.Ns, wi) < 0;
                const f3 orig = rayInside ? s.pos - s.Ns * T_EPSILON : s.pos + s.Ns * T_EPSILON;
                cont = true;
                no = make_float4(orig.x, orig.y, orig.z, __uint_as_float(i));
                nd = make_float4(wi.x, wi.y, wi.z, __uint_as_float((kind << 8) | (k + 1u)));
                ntp = make_float4(tp.x, tp.y, tp.z, 0.f);
              }
            }
          }
        }
      }
    }
    const unsigned at = warp_append(&b.ctl->n_next, cont);
    if (cont) {
      __stcs(b.q_o[nxt] + at, no);
      __stcs(b.q_d[nxt] + at, nd);
      __stcs(b.q_tp[nxt] + at, ntp);
    }
  }
}

// ---- MISweight, BDPT.hpp:70-222 ----------------------------------------------------------------
// sEnd / tEnd are lpverts[s-1] / epverts[t-1] (ignored when s == 0 / t == 1 respectively).
static __device__ __noinline__ float bdpt_mis_weight(const DevScene& sc, const BdptCam& cam, const BdptBuffers& b, unsigned i,
                                              int s, int t, const Vtx sEnd, const Vtx tEnd, float camFwd) {
  if (s + t == 2) return 1.f;
  float pdf_tEndFwd = 0.f, pdf_tEndRev = 0.f, pdf_sEndFwd = 0.f, pdf_sEndRev = 0.f, G_connect = 0.f;
  if (s == 0) {
    const f3 wo = normalized(eye_pos(cam, b, i, t - 2) - tEnd.pos);
    const float c = fabsf(dot(tEnd.Ng, wo));
    float dirpdf = c / T_PI;
    dirpdf = dirpdf / c;
    const bool sphere = (tEnd.code & kSphereBit) != 0u;
    const float pickpdf = (sc.n_lights > 0 && tEnd.m.has_emission)
                              ? 1.f / (sc.n_lights * slot_area(sc, tEnd.code & kSlotMask, sphere))
                              : 0.f;  // getLightPdf
    pdf_tEndFwd = pickpdf;
    pdf_tEndRev = dirpdf;
  } else {
    const f3 tpos = t == 1 ? ld3(cam.eye) : tEnd.pos;
    const f3 tNg = t == 1 ? ld3(cam.fwd) : tEnd.Ng;
    G_connect = Geo(sEnd.pos, sEnd.Ng, tpos, tNg);
    if (t == 1) {
      const f3 cam2sEnd = normalized(sEnd.pos - tpos);
      const float camcos = dot(tNg, cam2sEnd);
      const float d = cam.imagePlaneDist / camcos;
      pdf_tEndFwd = (cam.filmPlaneAreaInv * d * d / camcos) / camcos;
      pdf_tEndRev = cam.lensAreaInv;
      const f3 s2prev = normalized(light_pos(b, i, s - 2) - sEnd.pos);
      pdf_sEndFwd = mat_pdf_eval(sEnd.m, -cam2sEnd, s2prev, sEnd.Ns, sc.eta, sEnd.m.eta) / fabsf(dot(-cam2sEnd, sEnd.Ng));
      pdf_sEndRev = mat_pdf_eval(sEnd.m, s2prev, -cam2sEnd, sEnd.Ns, sc.eta, sEnd.m.eta) / fabsf(dot(s2prev, sEnd.Ng));
    } else if (s == 1) {
      const f3 light2tEnd = normalized(tEnd.pos - sEnd.pos);
      const float c = dot(sEnd.Ng, light2tEnd);
      pdf_sEndFwd = c / T_PI / c;
      pdf_sEndRev = light_mis(b, i, 0).rev;
      const f3 t2prev = normalized(eye_pos(cam, b, i, t - 2) - tEnd.pos);
      pdf_tEndFwd = mat_pdf_eval(tEnd.m, -light2tEnd, t2prev, tEnd.Ns, sc.eta, tEnd.m.eta) / fabsf(dot(-light2tEnd, tEnd.Ng));
      pdf_tEndRev = mat_pdf_eval(tEnd.m, t2prev, -light2tEnd, tEnd.Ns, sc.eta, tEnd.m.eta) / fabsf(dot(t2prev, tEnd.Ng));
    } else {
      const f3 s2t = normalized(tEnd.pos - sEnd.pos);
      const f3 s2prev = normalized(light_pos(b, i, s - 2) - sEnd.pos);
      const f3 t2prev = normalized(eye_pos(cam, b, i, t - 2) - tEnd.pos);
      pdf_sEndFwd = mat_pdf_eval(sEnd.m, s2t, s2prev, sEnd.Ns, sc.eta, sEnd.m.eta) / fabsf(dot(s2t, sEnd.Ng));
      pdf_sEndRev = mat_pdf_eval(sEnd.m, s2prev, s2t, sEnd.Ns, sc.eta, sEnd.m.eta) / fabsf(dot(s2prev, sEnd.Ng));
      pdf_tEndFwd = mat_pdf_eval(tEnd.m, -s2t, t2prev, tEnd.Ns, sc.eta, tEnd.m.eta) / fabsf(dot(-s2t, tEnd.Ng));
      pdf_tEndRev = mat_pdf_eval(tEnd.m, t2prev, -s2t, tEnd.Ns, sc.eta, tEnd.m.eta) / fabsf(dot(t2prev, tEnd.Ng));
    }
  }
  // misnodes 0..k, light end first (:147-184)
  float toLight[kBdptMaxLen + 2], toEye[kBdptMaxLen + 2];
  unsigned deltaMask = 0u;
  const int k = s + t - 1;
#pragma unroll 1
  for (int n = 0; n < s - 1; ++n) {
    const MisRec a = light_mis(b, i, n), nx = light_mis(b, i, n + 1);
    toLight[n] = (n == 0) ? a.rev : a.rev * a.G;
    toEye[n] = a.fwd * nx.G;
    if (a.delta) deltaMask |= 1u << n;
  }
  if (s > 0) {
    const MisRec a = light_mis(b, i, s - 1);
    toLight[s - 1] = (s == 1) ? pdf_sEndRev : pdf_sEndRev * a.G;
    toEye[s - 1] = pdf_sEndFwd * G_connect;
    if (a.delta) deltaMask |= 1u << (s - 1);
  }
#pragma unroll 1
  for (int ti = 0; ti < t - 1; ++ti) {
    const MisRec a = eye_mis(cam, b, i, ti, camFwd), nx = eye_mis(cam, b, i, ti + 1, camFwd);
    toEye[k - ti] = (ti == 0) ? a.rev : a.rev * a.G;
    toLight[k - ti] = a.fwd * nx.G;
    if (a.delta) deltaMask |= 1u << (k - ti);
  }
  {
    const MisRec a = eye_mis(cam, b, i, t - 1, camFwd);
    toEye[k - (t - 1)] = (t == 1) ? pdf_tEndRev : pdf_tEndRev * a.G;
    toLight[k - (t - 1)] = (s == 0) ? pdf_tEndFwd : pdf_tEndFwd * G_connect;
    if (a.delta) deltaMask |= 1u << (k - (t - 1));
  }
  auto isDelta = [&](int n) { return ((deltaMask >> n) & 1u) != 0u; };
  float p_i_plus_1 = 1.0f, denominator = 1.0f;
#pragma unroll 1
  for (int n = s; n < k; ++n) {
    if (n == 0) {
      p_i_plus_1 *= toLight[0] / toLight[1];
      if (isDelta(1)) continue;
    } else {
      p_i_plus_1 *= toEye[n - 1] / toLight[n + 1];
      if (isDelta(n) || isDelta(n + 1)) continue;
    }
    denominator += p_i_plus_1 * p_i_plus_1;
  }
  float p_i_minus_1 = 1.0f;
#pragma unroll 1
  for (int n = s; n > 0; --n) {
    if (n == 1) {
      p_i_minus_1 *= toLight[1] / toLight[0];
      if (isDelta(0)) continue;
    } else {
      p_i_minus_1 *= toLight[n] / toEye[n - 2];
      if (isDelta(n - 1) || isDelta(n - 2)) continue;
    }
    denominator += p_i_minus_1 * p_i_minus_1;
  }
  const float res = 1 / denominator;
  if (res < T_MIN_DIVISOR || isnan(res) || isinf(res)) return 0.f;
  return res;
}

// ---- connections of one path length, BDPT.hpp:752-886 --------------------------------------------
// thread = (s, sample): idx = s * n + i, t = pathLength + 1 - s; a warp holds 32 samples of one strategy
__global__ void __launch_bounds__(256, 2)
bdpt_connect(const __grid_constant__ DevScene sc, const __grid_constant__ BdptCam cam, BdptBuffers b, int pathLength,
             unsigned long long first_path, unsigned n) {
  const unsigned npix = (unsigned)cam.width * (unsigned)cam.height;
  const unsigned long long total = (unsigned long long)(pathLength + 1) * n;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < total;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long idx = base + threadIdx.x;
    bool want = false;
    float4 so = make_float4(0, 0, 0, 0), sd = so, scn = so;
    if (idx < total) {
      const int s = (int)(idx / n);
      const unsigned i = (unsigned)(idx % n);
      const int t = pathLength + 1 - s;
      const int nE = b.nE[i], nL = b.nL[i];
      // epverts.size() < 2 -> continue (:750); t > epverts.size() || s > lpverts.size() -> continue (:758)
      if (nE >= 2 && t <= nE && s <= nL) {
        const uint32_t pixel = (uint32_t)((first_path + i) % npix);
        f3 pixelPos, rayDir;
        bdpt_pixel(cam, pixel, pixelPos, rayDir);
        float camFwd, tp0;
        bdpt_cam_pdfs(cam, pixelPos, rayDir, camFwd, tp0);
        Vtx none;
        none.pos = none.Ng = none.Ns = none.tp = mk(0.f);
        none.m = Mat{};
        none.code = 0u;
        if (s == 0) {  // the eye path hit a light, :766-785
          const Vtx e = load_vtx(sc, b, t - 2, i);
          if (e.m.emission.x || e.m.emission.y || e.m.emission.z) {
            int dummy;
            const float we = We(cam, pixelPos, &dummy);
            const f3 contrib = we * e.tp * e.m.emission;
            if (norm2(contrib) != 0 && !isnan(contrib.x)) {
              const float misw = bdpt_mis_weight(sc, cam, b, i, s, t, none, e, camFwd);
              const f3 c = misw * contrib;
              float* p = b.accum + (size_t)pixel * 3;
              atomicAdd(p + 0, c.x);
              atomicAdd(p + 1, c.y);
              atomicAdd(p + 2, c.z);
            }
          }
        } else if (t == 1) {  // light vertex -> camera, :788-834
          const Vtx lv = load_vtx(sc, b, kBdptEyeSlots + s - 1, i);
          if (!(lv.m.emission.x || lv.m.emission.y || lv.m.emission.z)) {
            const Vtx l0 = load_vtx(sc, b, kBdptEyeSlots + 0, i);
            const f3 l = l0.m.emission;
            const f3 eye = ld3(cam.eye), fwd = ld3(cam.fwd);
            const f3 wi = normalized(eye - lv.pos);
            f3 bsdf = mk(1.f);
            bool rayInside = false;
            if (s != 1) {
              const f3 wo = normalized(light_pos(b, i, s - 2) - lv.pos);
              rayInside = dot(wi, lv.Ng) < 0;  // sic: Ng (BDPT.hpp:803)
              bsdf = BxDF_adjoint(lv.m, wi, wo, lv.Ng, lv.Ns, sc.eta);
            }
            const float G = Geo(eye, fwd, lv.pos, lv.Ng);
            int index;
            const float we = We(cam, lv.pos, &index);
            const f3 contrib = l * bsdf * lv.tp * G * we;  // x SPP_inv in bdpt_finalize
            if (norm2(contrib) != 0 && !isnan(contrib.x)) {
              const float misw = bdpt_mis_weight(sc, cam, b, i, s, t, lv, none, camFwd);
              if (misw != 0.f && dot(wi, fwd) < 0) {
                const f3 orig = rayInside ? lv.pos - lv.Ns * T_EPSILON : lv.pos + lv.Ns * T_EPSILON;
                const f3 dv = eye - orig;
                const f3 dir = normalized(dv);
                const f3 c = misw * contrib;
                want = true;
                so = make_float4(orig.x, orig.y, orig.z, sqrtf(norm2(dv)));
                sd = make_float4(dir.x, dir.y, dir.z, __uint_as_float((uint32_t)index));
                scn = make_float4(c.x, c.y, c.z, 0.f);
              }
            }
          }
        } else {  // general connection, :836-885
          const Vtx e = load_vtx(sc, b, t - 2, i);
          if (!(e.m.emission.x || e.m.emission.y || e.m.emission.z)) {
            const Vtx lv = load_vtx(sc, b, kBdptEyeSlots + s - 1, i);
            const f3 l = s == 1 ? lv.m.emission : load_vtx(sc, b, kBdptEyeSlots + 0, i).m.emission;
            const f3 connectDir = normalized(e.pos - lv.pos);
            const f3 e_wo = normalized(eye_pos(cam, b, i, t - 2) - e.pos);
            const f3 evBSDF = BxDF(e.m, -connectDir, e_wo, e.Ng, e.Ns, sc.eta);
            f3 lvBSDF, l_wo = mk(0.f);
            if (s == 1) {
              lvBSDF = dot(connectDir, lv.Ns) >= 0 ? mk(1.f) : mk(0.f);
            } else {
              l_wo = normalized(light_pos(b, i, s - 2) - lv.pos);
              lvBSDF = BxDF_adjoint(lv.m, connectDir, l_wo, lv.Ng, lv.Ns, sc.eta);
            }
            int dummy;
            const float we = We(cam, pixelPos, &dummy);
            const float G = Geo(e.pos, e.Ng, lv.pos, lv.Ng);
            const f3 contrib = we * e.tp * evBSDF * G * lv.tp * lvBSDF * l;
            if (norm2(contrib) != 0 && !isnan(contrib.x)) {
              const float misw = bdpt_mis_weight(sc, cam, b, i, s, t, lv, e, camFwd);
              if (misw != 0.f) {
                const bool eInside = dot(e_wo, e.Ns) < 0;
                const f3 eOrig = eInside ? e.pos - e.Ns * T_EPSILON : e.pos + e.Ns * T_EPSILON;
                const bool lInside = s == 1 ? false : dot(l_wo, lv.Ns) < 0;
                const f3 lorig = lInside ? lv.pos - lv.Ns * T_EPSILON : lv.pos + lv.Ns * T_EPSILON;
                const f3 dv = lorig - eOrig;
                const f3 dir = normalized(dv);
                const f3 c = misw * contrib;
                want = true;
                so = make_float4(eOrig.x, eOrig.y, eOrig.z, sqrtf(norm2(dv)));
                sd = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pixel));
                scn = make_float4(c.x, c.y, c.z, 0.f);
              }
            }
          }
        }
      }
    }
    const unsigned at = warp_append(&b.ctl->n_shadow, want);
    if (want) {
      b.sh_o[at] = so;
      b.sh_d[at] = sd;
      b.sh_c[at] = scn;
    }
  }
}

// ---- finalize: FrameBuffer starts at bkgcolor (Camera.hpp:28), contributions are added ---------
__global__ void bdpt_finalize(const float* __restrict__ accum, float inv_spp, float b0, float b1, float b2,
                              float* __restrict__ out, size_t npix) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    out[3 * i + 0] = b0 + accum[3 * i + 0] * inv_spp;
    out[3 * i + 1] = b1 + accum[3 * i + 1] * inv_spp;
    out[3 * i + 2] = b2 + accum[3 * i + 2] * inv_spp;
  }
}

}  // namespace tutu
