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
// appended through warp-aggregated atomics (one atomicAdd per warp).  Free slots are refilled by
// raygen every iteration, so the wavefront stays full until the last samples.
#pragma once
#include <cstddef>
#include "shade.cuh"

namespace tutu {

constexpr uint32_t kModeFresh = 0;   // ray traced by traceRay itself: miss -> bkgcolor
constexpr uint32_t kModeXInter = 1;  // ray traced as x_inter of the previous vertex
constexpr uint32_t kFlagMirror = 1u << 9;  // previous vertex PERFECT_REFLECTIVE (PathTracing.hpp:252)
constexpr uint32_t kShadowFinal = 0xFFFFFFFFu;

struct RayGenK {
  float eye[3], ul[3], dh[3], dv[3], coh[3], cov[3];
  int width, height;
};

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

// ---- control ---------------------------------------------------------------------------------
__global__ void wf_ctl_after_raygen(WfCtl* ctl, unsigned capacity) {
  const unsigned free_slots = capacity - ctl->n_cur;
  const unsigned long long left = ctl->total_paths - ctl->next_path;
  const unsigned add = (unsigned)(left < (unsigned long long)free_slots ? left : free_slots);
  ctl->n_cur += add;
  ctl->next_path += add;
  ctl->n_next = 0;
  ctl->n_shadow = 0;
  ctl->cursor_extend = 0;
  ctl->cursor_shadow = 0;
  ctl->sum_extend += ctl->n_cur;
  for (int c = 0; c < 8; ++c) ctl->class_count[c] = 0;
}
__global__ void wf_ctl_after_iter(WfCtl* ctl) {
  ctl->sum_shadow += ctl->n_shadow;
  ctl->n_cur = ctl->n_next;
  ctl->iterations += 1;
  ctl->done = (ctl->n_cur == 0 && ctl->next_path >= ctl->total_paths) ? 1u : 0u;
}

// ---- raygen: PathTracing.hpp:499-509 ---------------------------------------------------------
// path g -> pixel = g % npix, sample = sample_begin + g / npix; primary rays carry no jitter, the
// pixel position is ul + x*delta_h + y*delta_v + c_off_v + c_off_v (sic), evaluated with the
// reference's operation order in exact fp32.
__global__ void __launch_bounds__(256)
wf_raygen(WfBuffers b, int cur, RayGenK k, unsigned sample_begin) {
  const WfCtl c = *b.ctl;
  const unsigned free_slots = b.capacity - c.n_cur;
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
constexpr unsigned kPacketRays = 128;  // rays per queue fetch (one same-address atomic each)

// Ray packets pulled from the queue with one atomicAdd per warp (lane 0) and a shuffle
// broadcast: a warp that drew short rays moves on to the next packet instead of idling behind the
// slowest warp of a statically partitioned grid.
template <bool SMALL>
__global__ void __launch_bounds__(256)
wf_extend(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int cur) {
  extern __shared__ unsigned long long s_stack[];  // !SMALL: traversal stack (trace.cuh: SharedStack)
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
        if (SMALL)
          traverse_small<false>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h);
        else  // incoherent queue: per-lane walk (batched primitive tests only pay on sorted batches)
          traverse_shared<false>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, 0.f, h, s_stack);
        __stcs(b.hit + i, make_float4(h.t, h.u, h.v, __int_as_float(h.slot)));
      }
      __syncwarp();
    }
  }
}

// ---- shadow: isShadowRayBlocked -> hasIntersection, then the deferred NEE add ------------------
template <bool SMALL>
__global__ void __launch_bounds__(256)
wf_shadow(const __grid_constant__ DevScene sc, const __grid_constant__ SmallScene ss, WfBuffers b, int nxt) {
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
      if (j < n) {
        const float4 o = __ldcs(b.sh_o + j);
        const float4 d = __ldcs(b.sh_d + j);
        Hit h;
        const bool blocked = SMALL ? traverse_small<true>(sc, ss, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h)
                                   : traverse_variant<true, 3>(sc, Ray{o.x, o.y, o.z, d.x, d.y, d.z}, o.w, h);
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

// ---- shade -------------------------------------------------------------------------------------
struct Surf {  // Intersection (Intersection.hpp:13-31) rebuilt from the 16-byte hit record
  f3 pos, Ng, Ns;
  float tu, tv;
  Mat m;
  bool textured;
  bool sphere;
  uint32_t slot;
};

__device__ __forceinline__ Surf load_surface(const DevScene& sc, const Ray& r, const float4 hit) {
  Surf s;
  const uint32_t code = __float_as_uint(hit.w);
  s.slot = code & kSlotMask;
  s.sphere = (code & kSphereBit) != 0u;
  const float t = hit.x;
  s.pos = mk(r.ox, r.oy, r.oz) + t * mk(r.dx, r.dy, r.dz);  // Triangle.hpp:54
  const float4* sh = sc.shade + 4 * (size_t)s.slot;
  const float4 s3 = __ldg(sh + 3);
  const uint32_t flags = __float_as_uint(s3.w);
  s.textured = (flags & 0x80000000u) != 0u;
  s.m = load_material(sc, (int)(flags & 0x3FFFFFFFu));
  s.tu = s.tv = 0.f;
  if (s.sphere) {
    const float4 g0 = __ldg(sc.geom + 3 * (size_t)s.slot);
    s.Ng = normalized(s.pos - mk(g0.x, g0.y, g0.z));  // Sphere.hpp:54-55
    s.Ns = s.Ng;
    if (s.textured) {  // Sphere.hpp:58-72
      float phi = acosf(s.Ng.z);
      s.tv = phi / T_PI;
      float theta = atan2f(s.Ng.y, s.Ng.x);
      if (theta < 0) theta += 2 * T_PI;
      s.tu = theta / (2.f * T_PI);
    }
  } else {
    const float4 g2 = __ldg(sc.geom + 3 * (size_t)s.slot + 2);
    s.Ng = mk(g2.y, g2.z, g2.w);
    const float4 s0 = __ldg(sh + 0), s1 = __ldg(sh + 1), s2 = __ldg(sh + 2);
    const float u = hit.y, v = hit.z, w = 1 - u - v;
    // Triangle.hpp:56
    s.Ns = normalized(mk(s0.x, s0.y, s0.z) * w + mk(s1.x, s1.y, s1.z) * u + mk(s2.x, s2.y, s2.z) * v);
    if (s.textured) {  // Triangle.hpp:62-69
      s.tu = s0.w * w + s2.w * u + s3.y * v;
      s.tv = s1.w * w + s3.x * u + s3.z * v;
    }
  }
  return s;
}

// textureModify + changeNormalDir, IIntegrator.hpp:27-127.  By value in and out (a reference to the
// caller's Surf would pin that whole record to local memory on the untextured path too).
struct TexMod {
  f3 diffuse, Ns;
  float roughness, metallic;
};
__device__ __noinline__ TexMod texture_modify(const DevScene& sc, uint32_t slot, bool sphere, float tu, float tv,
                                              f3 Ng, TexMod in) {
  TexMod r = in;
  const int4 ti = __ldg(sc.leaftex + slot);
  if (ti.x != -1) r.diffuse = tex_fetch(sc, 0, ti.x, tu, tv);
  if (ti.y != -1) {
    const f3 color = tex_fetch(sc, 1, ti.y, tu, tv);
    f3 T, B, nDir;
    if (!sphere) {
      const float4* g = sc.geom + 3 * (size_t)slot;
      const float4 a = __ldg(g + 0), b = __ldg(g + 1), c = __ldg(g + 2);
      const f3 e1 = mk(a.w, b.x, b.y), e2 = mk(b.z, b.w, c.x);
      const float4* sh = sc.shade + 4 * (size_t)slot;
      const float4 s0 = __ldg(sh + 0), s1 = __ldg(sh + 1), s2 = __ldg(sh + 2), s3 = __ldg(sh + 3);
      nDir = normalized(in.Ns);
      const float deltaU1 = s2.w - s0.w, deltaV1 = s3.x - s1.w;
      const float deltaU2 = s3.y - s0.w, deltaV2 = s3.z - s1.w;
      const float coef = 1 / (-deltaU1 * deltaV2 + deltaV1 * deltaU2);
      T = normalized(coef * (-deltaV2 * e1 + deltaV1 * e2));
      B = normalized(coef * (-deltaU2 * e1 + deltaU1 * e2));
    } else {
      nDir = Ng;
      const float q = sqrtf(nDir.x * nDir.x + nDir.y * nDir.y);
      T = mk(-nDir.y / q, nDir.x / q, 0.f);
      B = cross(nDir, T);
    }
    f3 res;
    res.x = T.x * color.x + B.x * color.y + nDir.x * color.z;
    res.y = T.y * color.x + B.y * color.y + nDir.y * color.z;
    res.z = T.z * color.x + B.z * color.y + nDir.z * color.z;
    r.Ns = normalized(res);
  }
  if (ti.z != -1) r.roughness = tex_fetch(sc, 2, ti.z, tu, tv).x;
  if (ti.w != -1) r.metallic = tex_fetch(sc, 3, ti.w, tu, tv).x;
  return r;
}

// Object::getArea of the primitive in a leaf slot (getLightPdf, IIntegrator.hpp:155-168)
__device__ __forceinline__ float slot_area(const DevScene& sc, uint32_t slot, bool sphere) {
  const float4* g = sc.geom + 3 * (size_t)slot;
  const float4 a = __ldg(g + 0);
  if (sphere) return a.w * a.w * T_PI;
  const float4 b = __ldg(g + 1), c = __ldg(g + 2);
  const f3 cr = cross(mk(a.w, b.x, b.y), mk(b.z, b.w, c.x));
  return sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z) * 0.5f;
}

struct LightSample {
  f3 pos, Ns, emission;
  float pdf;  // 1 / (size * area), IIntegrator.hpp:191
};

// sampleLight + Triangle/Sphere::samplePoint
__device__ __forceinline__ LightSample sample_light(const DevScene& sc, float r_idx, float ra, float rb) {
  const int size = sc.n_lights;
  int index = (int)(r_idx * (size - 1) + 0.4999f);  // IIntegrator.hpp:184 (sic, non-uniform)
  if (size == 1) index = 0;
  const float4* L = sc.lights + 8 * (size_t)index;
  const float4 l0 = __ldg(L + 0), l1 = __ldg(L + 1), l2 = __ldg(L + 2), l3 = __ldg(L + 3);
  const float4 l4 = __ldg(L + 4), l5 = __ldg(L + 5), l6 = __ldg(L + 6);
  LightSample s;
  s.emission = mk(l6.x, l6.y, l6.z);
  const float area = l0.w;
  if (__float_as_int(l1.w) == TUTU_PRIM_SPHERE) {  // Sphere.hpp:139-164
    const float radius = l3.w;
    const float theta = ra * 2 * T_PI;
    const float phi = rb * T_PI;
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    const f3 c = mk(l0.x, l0.y, l0.z);
    s.pos = mk(c.x + radius * ct * sp, c.y + radius * st * sp, c.z + radius * cp);
    s.Ns = normalized(s.pos - c);
  } else {  // Triangle.hpp:119-142
    const float u = ra;
    const float v = rb * (1 - u);
    const float w = 1 - u - v;
    s.pos = w * mk(l0.x, l0.y, l0.z) + u * mk(l1.x, l1.y, l1.z) + v * mk(l2.x, l2.y, l2.z);
    s.Ns = normalized(w * mk(l3.x, l3.y, l3.z) + u * mk(l4.x, l4.y, l4.z) + v * mk(l5.x, l5.y, l5.z));
  }
  s.pdf = fdiv(1.f, size * area);
  return s;
}

// the rare in-kernel shadow ray of PathTracing.hpp:215 (kept out of line: it owns a traversal stack)
__device__ __noinline__ bool shadow_blocked_inline(const DevScene& sc, const Ray r, const float dist) {
  Hit h;
  return traverse<true, 0, false>(sc, r, dist, h, nullptr);
}

struct ShadeOut {
  bool cont;      // a continuation ray goes to the next queue
  bool shadow;    // an NEE shadow ray goes to the shadow queue
  bool finished;  // the path ended at this vertex (L must reach the frame buffer)
  // continuation
  f3 o, d, beta, tp, fcos;
  float q;  // 2 (o - x) . d for the next vertex's r^2
  float mat_pdf;
  float rr_u;
  uint32_t depth_mode;
  // shadow
  f3 so, sd, sc;
  float sdist;
};

// SPEC bit 0: every material of the scene is LAMBERTIAN; bit 1: no primitive is textured — compile-time
// removal of unreachable material code.  Measured: the all-Lambertian kernel still
// needs > 80 registers (216 B of spills at 3 blocks/SM), so it buys no occupancy and only SPEC = 0 is
// instantiated; the hook is kept for scenes where the general kernel's size matters.
constexpr int kSpecLambertOnly = 1, kSpecNoTextures = 2;
template <int SPEC>
__device__ __forceinline__ void shade_vertex(const DevScene& sc, uint64_t seed, const Ray& ray,
                                             const float4 hit, uint32_t pixel, uint32_t sample,
                                             uint32_t depth, uint32_t mode, uint32_t flags, f3 beta,
                                             f3 tp, f3& L, const float4 st3, float q_prev, float rr_u,
                                             ShadeOut& out) {
  out.cont = out.shadow = false;
  out.finished = true;
  const int slotcode = __float_as_int(hit.w);
  if (slotcode < 0) {
    // PathTracing.hpp:150: only a ray traced by traceRay itself sees the background;
    // a missed x_inter (:234) just ends the path.
    if (mode == kModeFresh) L = L + beta * mk(sc.bkg[0], sc.bkg[1], sc.bkg[2]);
    return;
  }
  Surf s = load_surface(sc, ray, hit);
  const bool kLamb = (SPEC & kSpecLambertOnly) != 0, kNoTex = (SPEC & kSpecNoTextures) != 0;
  const f3 dir = mk(ray.dx, ray.dy, ray.dz);

  if (mode == kModeXInter) {
    // ---- second half of the previous vertex, PathTracing.hpp:236-278 ----
    const f3 fcos = mk(st3.x, st3.y, st3.z);  // f_r * cos_theta
    const float mat_pdf = st3.w;
    float light_pdf = 0.f;
    if (s.m.has_emission && sc.n_lights > 0) light_pdf = fdiv(1.f, sc.n_lights * slot_area(sc, s.slot, s.sphere));
    bool as_light = false;
    if (light_pdf) {
      const f3 light_N = normalized(s.Ns);
      const float cos_theta_prime = dot(light_N, -dir);
      if (cos_theta_prime > 0) {
        as_light = true;
        // r2 = |x_inter.pos - inter.pos|^2 (PathTracing.hpp:246) with x_inter.pos = o + t d and
        // o = inter.pos -+ EPSILON Ns: |o - x|^2 + t^2 |d|^2 + 2 t (o - x).d.  The previous vertex
        // passes q = 2 (o - x).d along with the ray instead of its position (16 B less per vertex in
        // each direction); |o - x| = EPSILON |Ns| and |d| are 1 to rounding.
        const float t_hit = hit.x;
        const float r2 = T_EPSILON * T_EPSILON + t_hit * t_hit + q_prev * t_hit;
        const float l_pdf_transformed = fdiv(light_pdf * r2, cos_theta_prime);
        float mis_weight_m = getMisWeight(mat_pdf, l_pdf_transformed);
        if ((flags & kFlagMirror) && mat_pdf == 1.f) mis_weight_m = 1.f;
        if (mat_pdf < T_MIN_DIVISOR) return;
        L = L + beta * (mis_weight_m * s.m.emission * fcos / mat_pdf);
        return;
      }
    }
    if (!as_light) {
      // jmp2: Russian roulette on tp, reset while depth <= MIN_DEPTH (:265-273)
      if (!(depth > T_MIN_DEPTH)) tp = mk(1.f);
      const float rr_prob = max3(tp);
      if (rr_u > rr_prob) return;  // rr_u = slot 5 of this depth's Philox stream, drawn with the BSDF sample
      const f3 coe = fcos / (mat_pdf * rr_prob);
      if (mat_pdf * rr_prob < T_MIN_DIVISOR) return;
      tp = tp * coe;
      beta = beta * coe;
      depth += 1;
      if (depth > T_MAX_DEPTH) {  // traceRay(depth+1) returns 0 (:140)
        L = L + beta * 0.f;
        return;
      }
    }
  }

  // ---- traceRay body at `depth` with inter = this hit (:152-232) ----
  const f3 wo = -dir;
  if (!kLamb && (s.m.type == TUTU_MAT_PERFECT_REFRACTIVE || s.m.type == TUTU_MAT_MICROFACET_T)) {
    // calcForRefractive (:80-134): no textures, no NEE, no roulette
    const Rand6 rn = draw6(seed, pixel, sample, depth);
    float eta_i = sc.eta, eta_t = s.m.eta;
    f3 wi = mk(0.f);
    const int ok = sampleDirection(s.m, wo, s.Ns, wi, eta_i, rn.u[3], rn.u[4], rn.u[5]);
    const bool TIR = (ok & 2) != 0;
    wi = normalized(wi);
    float pdf = mat_pdf_eval(s.m, wi, wo, s.Ns, eta_i, eta_t);
    if (TIR) {
      wi = normalized(getReflectionDir(wo, s.Ns));
      pdf = 1;
      if (s.m.type == TUTU_MAT_MICROFACET_T) {
        f3 interNs = s.Ns;
        if (dot(wo, s.Ng) < 0) {
          const float sw = eta_i;
          eta_i = eta_t;
          eta_t = sw;
          interNs = -interNs;
        }
        const f3 h = normalized(wo + wi);
        const float cosTheta = fabsf(dot(interNs, h));
        wi = normalized(getReflectionDir(wo, h));
        pdf = 1 * D_ndf(h, interNs, s.m.roughness) * cosTheta / (4.f * dot(wo, h));
      }
    }
    const f3 f_r = BxDF(s.m, wi, wo, s.Ng, s.Ns, eta_i, TIR);
    f3 rayOrig = s.pos;
    float cosv;
    if (dot(wi, s.Ns) > 0) {
      rayOrig = rayOrig + s.Ns * T_EPSILON;
      cosv = fabsf(dot(s.Ng, wi));
    } else {
      rayOrig = rayOrig - s.Ns * T_EPSILON;
      cosv = fabsf(dot(-s.Ng, wi));
    }
    // the reference recurses first and tests pdf afterwards (:128-133); testing first is equivalent
    if (pdf < T_MIN_DIVISOR) return;
    beta = beta * (cosv * f_r / pdf);
    if (depth + 1 > T_MAX_DEPTH) {
      L = L + beta * 0.f;
      return;
    }
    out.cont = true;
    out.finished = false;
    out.o = rayOrig;
    out.d = wi;
    out.beta = beta;
    out.tp = mk(1.f);
    out.fcos = mk(0.f);
    out.mat_pdf = 0.f;
    out.rr_u = 0.f;
    out.q = 0.f;
    out.depth_mode = (depth + 1) | (kModeFresh << 8);
    return;
  }

  if (!kNoTex && s.textured) {
    const TexMod tm = texture_modify(sc, s.slot, s.sphere, s.tu, s.tv, s.Ng,
                                     TexMod{s.m.diffuse, s.Ns, s.m.roughness, s.m.metallic});
    s.m.diffuse = tm.diffuse;
    s.Ns = tm.Ns;
    s.m.roughness = tm.roughness;
    s.m.metallic = tm.metallic;
  }
  if (!kLamb && s.m.type == TUTU_MAT_UNLIT) {  // :161
    L = L + beta * s.m.diffuse;
    return;
  }
  const bool emissive = s.m.emission.x || s.m.emission.y || s.m.emission.z;
  if (emissive) {  // :164-170
    L = L + beta * (depth > 0 ? mk(0.f) : s.m.emission);
    return;
  }

  const Rand6 rn = draw6(seed, pixel, sample, depth);

  // ---- NEE, :185-218 ----
  if (sc.n_lights > 0) {
    const LightSample ls = sample_light(sc, rn.u[0], rn.u[1], rn.u[2]);
    const bool rayInside = dot(s.Ns, wo) < 0;
    const f3 shadowRayOrig = rayInside ? s.pos - s.Ns * T_EPSILON : s.pos + s.Ns * T_EPSILON;
    const f3 lightPos = ls.pos + ls.Ns * T_EPSILON;
    f3 wi = ls.pos - s.pos;
    const float r2 = norm2(wi);
    wi = normalized(wi);
    if (!(dot(wi, ls.Ns) > 0)) {
      const float mat_pdf = mat_pdf_eval(s.m, wi, wo, s.Ns, sc.eta, s.m.eta);
      const f3 light_N = normalized(ls.Ns);
      const float cos_theta_prime = dot(light_N, -wi);
      if (cos_theta_prime > 0) {
        const float cos_theta = fabsf(dot(s.Ng, wi));
        const float pdfl = ls.pdf;
        const float light_pdf = fdiv(pdfl * r2, cos_theta_prime);
        const float mis_weight_l = getMisWeight(light_pdf, mat_pdf);
        const f3 f_r = BxDF(s.m, wi, wo, s.Ng, s.Ns, sc.eta);
        // isShadowRayBlocked (IIntegrator.hpp:135-153)
        const f3 sd = normalized(lightPos - shadowRayOrig);
        const f3 dv = lightPos - shadowRayOrig;
        const float dist = sqrtf(dv.x * dv.x + dv.y * dv.y + dv.z * dv.z);
        if (r2 * pdfl < T_MIN_DIVISOR) {
          // :215 — an unoccluded sample this close to the light ends the whole path; the
          // decision needs the visibility now, so this rare case traces its shadow ray inline.
          if (!shadow_blocked_inline(sc, Ray{shadowRayOrig.x, shadowRayOrig.y, shadowRayOrig.z, sd.x, sd.y, sd.z}, dist))
            return;
        } else {
          out.shadow = true;
          out.so = shadowRayOrig;
          out.sd = sd;
          out.sdist = dist;
          out.sc = beta * (mis_weight_l * ls.emission * f_r * cos_theta * cos_theta_prime / (r2 * pdfl));
        }
      }
    }
  }

  // ---- BSDF sample, :221-232 ----
  f3 wi = mk(0.f);
  const int ok = sampleDirection(s.m, wo, s.Ns, wi, sc.eta, rn.u[3], rn.u[4], rn.u[5]);
  if (!(ok & 1)) return;
  const float mat_pdf = mat_pdf_eval(s.m, wi, wo, s.Ns, sc.eta, s.m.eta);
  const bool inside = dot(wi, s.Ns) < 0;
  const f3 rayOrig = inside ? s.pos - s.Ns * T_EPSILON : s.pos + s.Ns * T_EPSILON;
  const float cos_theta = fabsf(dot(s.Ng, wi));
  const f3 f_r = BxDF(s.m, wi, wo, s.Ng, s.Ns, sc.eta);
  out.cont = true;
  out.finished = false;
  out.o = rayOrig;
  out.d = wi;
  out.beta = beta;
  out.tp = tp;
  out.fcos = f_r * cos_theta;
  out.mat_pdf = mat_pdf;
  out.rr_u = rn.u[5];
  out.q = 2.f * dot(rayOrig - s.pos, wi);
  out.depth_mode = depth | (kModeXInter << 8) | (s.m.type == TUTU_MAT_PERFECT_REFLECTIVE ? kFlagMirror : 0u);
}

// Compiled for up to 256 threads / 2 blocks per SM (128 registers); the launch picks the block size:
// a block waits at two barriers for one global atomic per iteration, so small blocks keep more
// independent groups in flight per SM, but many small blocks on different material code paths thrash
// the instruction cache.  Measured (Mpaths/s, 1024^2): Cornell 256x2 1606 / 128x4 1652 / 64x8 1670;
// glass + textures scene 426 / 353 / 317.  -> 64 threads for all-Lambertian untextured scenes, else 256.
#ifndef TUTU_SHADE_MIN_BLOCKS
#define TUTU_SHADE_MIN_BLOCKS 2
#endif
#ifndef TUTU_SHADE_BLOCK
#define TUTU_SHADE_BLOCK 256
#endif
constexpr int kShadeBlockSimple = 64;
template <int SPEC>
__device__ __forceinline__ void wf_shade_body(const DevScene& sc, const WfBuffers& b, int cur, uint64_t seed,
                                              unsigned (*s_cnt)[TUTU_SHADE_BLOCK / 32], unsigned* s_base, unsigned* s_pref) {
  const int nxt = cur ^ 1;
  const unsigned n = b.ctl->n_cur;
  // ncu (profiles/r01_shade_stalls.txt): with one atomicAdd per WARP on the two queue counters,
  // half of this kernel's stall samples sat on the shuffles waiting for those results — ~2 x 10^5
  // same-address atomics per launch serialise in one L2 slice.  The counts are therefore first
  // combined per BLOCK in shared memory (one global atomic per counter per block-iteration).
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // class lists (wf_classify): entry j of the concatenated lists -> queue index
  if (b.class_perm) {
    if (threadIdx.x == 0) {
      unsigned acc = 0u;
      for (int c = 0; c < kShadeClasses; ++c) {
        s_pref[c] = acc;
        acc += b.ctl->class_count[c];
      }
    }
    __syncthreads();
  }
  for (unsigned base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const unsigned j = base + threadIdx.x;
    const bool valid = j < n;
    unsigned i = j;
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
      // queue records are touched once per iteration: stream them past L1/L2 residency (.cs) so
      // the scene tables stay cached
      const float4 o = __ldcs(b.ray_o[cur] + i);
      const float4 d = __ldcs(b.ray_d[cur] + i);
      const float4 s0 = __ldcs(b.st0[cur] + i);
      s1 = __ldcs(b.st1[cur] + i);
      const float4 s2 = __ldcs(b.st2[cur] + i);
      const float4 hit = __ldcs(b.hit + i);
      const uint32_t dm = __float_as_uint(s2.w);
      const uint32_t depth = dm & 0xFFu, mode = (dm >> 8) & 1u;
      float4 s3 = make_float4(0, 0, 0, 0);
      if (mode == kModeXInter) s3 = __ldcs(b.st3[cur] + i);
      pixel = __float_as_uint(s0.w);
      L = mk(s2.x, s2.y, s2.z);
      Ray r{o.x, o.y, o.z, d.x, d.y, d.z};
      shade_vertex<SPEC>(sc, seed, r, hit, pixel, __float_as_uint(s1.w), depth, mode, dm, mk(s0.x, s0.y, s0.z),
                   mk(s1.x, s1.y, s1.z), L, s3, d.w, o.w, out);
    }
    // queue appends: warp ballots -> block prefix in shared memory -> one atomicAdd per counter
    const unsigned cmask = __ballot_sync(0xFFFFFFFFu, out.cont);
    const unsigned smask = __ballot_sync(0xFFFFFFFFu, out.shadow);
    if (lane == 0) {
      s_cnt[0][warp] = (unsigned)__popc(cmask);
      s_cnt[1][warp] = (unsigned)__popc(smask);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned tc = 0, ts = 0;
      const int n_warps = (int)(blockDim.x >> 5);
      for (int w = 0; w < n_warps; ++w) {
        const unsigned c = s_cnt[0][w], d = s_cnt[1][w];
        s_cnt[0][w] = tc, s_cnt[1][w] = ts;  // exclusive prefixes
        tc += c, ts += d;
      }
      unsigned long long base2 = 0ull;
      if (tc | ts)  // both queue counters in one atomic: {n_shadow : n_next}
        base2 = atomicAdd(reinterpret_cast<unsigned long long*>(&b.ctl->n_next), ((unsigned long long)ts << 32) | tc);
      s_base[0] = (unsigned)base2;
      s_base[1] = (unsigned)(base2 >> 32);
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    const unsigned ci = s_base[0] + s_cnt[0][warp] + (unsigned)__popc(cmask & lt);
    const unsigned si = s_base[1] + s_cnt[1][warp] + (unsigned)__popc(smask & lt);
    __syncthreads();  // s_cnt / s_base are rewritten by the next iteration
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
    } else if (valid && out.finished) {
      accum_add(b.accum, b.ctl, pixel, L);
    }
  }
}

__global__ void __launch_bounds__(TUTU_SHADE_BLOCK, TUTU_SHADE_MIN_BLOCKS)
wf_shade(const __grid_constant__ DevScene sc, WfBuffers b, int cur, uint64_t seed) {
  __shared__ unsigned s_cnt[2][TUTU_SHADE_BLOCK / 32];
  __shared__ unsigned s_base[2];
  __shared__ unsigned s_pref[kShadeClasses];
  wf_shade_body<0>(sc, b, cur, seed, s_cnt, s_base, s_pref);
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
