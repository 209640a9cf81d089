// vertex.cuh — one vertex of the PathTracing integrator (reference include/PathTracing.hpp:136-279
// traceRay MIS branch, :80-134 calcForRefractive): surface record, texture modifiers, light sampling and
// shade_vertex.  Device functions only; the kernels that call them are in wavefront.cuh (queues in HBM)
// and resident.cuh (path state in registers), which are compiled in separate translation units
// (ptxas 12.9 crashes on a unit that inlines shade_vertex into two kernels).
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
static __device__ __noinline__ TexMod texture_modify(const DevScene& sc, uint32_t slot, bool sphere, float tu, float tv,
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
static __device__ __noinline__ bool shadow_blocked_inline(const DevScene& sc, const Ray r, const float dist) {
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

}  // namespace tutu
