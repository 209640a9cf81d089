// shade.cuh — material evaluation / sampling, light sampling, texture fetch and the counter RNG
// for the wavefront path tracer.  Restates (citations relative to the reference's include/):
//   Material::BxDF / sampleDirection / pdf           Material.hpp:62-191, 200-343, 350-439
//   fresnel, fresnelSchlick, getReflectionDir, getRefractionDir, D_ndf, G_smf, getMisWeight,
//   offsetRayOrig, SphereLocal2world                 global.hpp:236-410
//   textureModify, changeNormalDir, Texture::getRGBat  IIntegrator.hpp:27-127, Texture.hpp:18-39
//   sampleLight, getLightPdf, Triangle/Sphere::samplePoint
//                                                    IIntegrator.hpp:155-192, Triangle.hpp:119-142,
//                                                    Sphere.hpp:139-164
// Shading arithmetic is ordinary fp32 (FMA contraction allowed, MUFU reciprocal / rsqrt with
// ~2 ulp instead of IEEE division / sqrt): it feeds a Monte Carlo estimate, whose parity with the
// reference is statistical.  Everything that selects a primitive lives in trace.cuh and is exact.
#pragma once
#include "trace.cuh"

namespace tutu {

#define T_PI 3.1415926535897f /* global.hpp:15 */
#define T_INV_PI (1.0f / 3.1415926535897f)
#define T_EPSILON 0.0005f     /* global.hpp:16 */
#define T_MIN_DIVISOR 0.04f   /* global.hpp:26 */
#define T_MAX_DEPTH 6         /* PathTracing.hpp:5 */
#define T_MIN_DEPTH 3         /* PathTracing.hpp:6 */

struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 mk(float a) { return f3{a, a, a}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float c) { return mk(a.x * c, a.y * c, a.z * c); }
__device__ __forceinline__ f3 operator*(float c, f3 a) { return mk(a.x * c, a.y * c, a.z * c); }
// a / b as a * MUFU.RCP(b): what __fdividef(a, b) computes for a normal b (same bits), without the three instructions
// that rescale a subnormal divisor (which this one flushes to zero: a / 0); 15 call sites in the Lambertian path.
__device__ __forceinline__ float fdiv(float a, float b) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
}
__device__ __forceinline__ f3 operator/(f3 a, float c) {
  const float r = fdiv(1.f, c);
  return mk(a.x * r, a.y * r, a.z * r);
}
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float norm2(f3 a) { return dot(a, a); }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// MUFU.RSQ alone.  rsqrtf() is the same instruction wrapped in six more that rescale subnormal arguments (ncu: 9 % of
// wf_shade's instructions over ~40 inlined normalisations, profiles/r02k_shade_lines.txt); for normal arguments the two
// return the same bits.
__device__ __forceinline__ float rsqrt_normal(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ f3 normalized(f3 v) {  // Vector.hpp:213-220
  const float l2 = v.x * v.x + v.y * v.y + v.z * v.z;
  // a select, not a branch: v * 1 is v (mag == 0 and NaN included; so is a vector shorter than 1.1e-19, whose squared
  // length is subnormal — the reference would still normalise it, no direction or normal of a scene gets there)
  const float inv = l2 >= 1.17549435e-38f ? rsqrt_normal(l2) : 1.f;
  return mk(v.x * inv, v.y * inv, v.z * inv);
}
__device__ __forceinline__ bool FLOAT_EQUAL(float x, float y) { return fabsf(x - y) < 0.0001f; }
__device__ __forceinline__ float max3(f3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
__device__ __forceinline__ bool any_nan(f3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }

// ---- counter RNG: Philox4x32-10, counter = (pixel, sample, depth, block), key = seed ---------
// slots per depth: 0 light index | 1,2 light point | 3,4 BSDF | 5 third BSDF draw or roulette
struct Rand6 {
  float u[6];
};
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                             uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
  c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// The six numbers of a path-tracing vertex: the six 21-bit fields of ONE Philox block (u = field * 2^-21,
// exact in fp32).  One block instead of two: Philox is ~13 % of wf_shade's instructions, and that kernel's
// time follows its instruction count.  The tests' CPU restatement draws the same fields.
__device__ __forceinline__ Rand6 draw6(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t depth) {
  Rand6 r;
  uint32_t a0 = pixel, a1 = sample, a2 = depth, a3 = 0u;
  philox4x32_10(a0, a1, a2, a3, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float s = 1.0f / 2097152.0f;
  r.u[0] = (float)(a0 >> 11) * s;
  r.u[1] = (float)(((a0 & 0x7FFu) << 10) | (a1 >> 22)) * s;
  r.u[2] = (float)((a1 >> 1) & 0x1FFFFFu) * s;
  r.u[3] = (float)(a2 >> 11) * s;
  r.u[4] = (float)(((a2 & 0x7FFu) << 10) | (a3 >> 22)) * s;
  r.u[5] = (float)((a3 >> 1) & 0x1FFFFFu) * s;
  return r;
}

// one Philox block: the four values of counter (pixel, sample, depth, block)
struct Rand4 {
  float u[4];
};
__device__ __forceinline__ Rand4 draw4(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t depth, uint32_t block) {
  uint32_t a0 = pixel, a1 = sample, a2 = depth, a3 = block;
  philox4x32_10(a0, a1, a2, a3, (uint32_t)seed, (uint32_t)(seed >> 32));
  Rand4 r;
  r.u[0] = u01(a0), r.u[1] = u01(a1), r.u[2] = u01(a2), r.u[3] = u01(a3);
  return r;
}

// ---- global.hpp helpers ----------------------------------------------------------------------
struct Mat {
  f3 diffuse;
  int type;
  f3 emission;
  float alpha, eta, roughness, metallic;
  int has_emission;
};

__device__ __forceinline__ Mat load_material(const DevScene& sc, int id) {
  const float4* m = sc.materials + 4 * (size_t)id;
  const float4 a = __ldg(m + 0), b = __ldg(m + 1), c = __ldg(m + 2), d = __ldg(m + 3);
  Mat r;
  r.diffuse = mk(a.x, a.y, a.z);
  r.type = __float_as_int(a.w);
  r.alpha = b.w;
  r.emission = mk(c.x, c.y, c.z);
  r.eta = c.w;
  r.roughness = d.x;
  r.metallic = d.y;
  r.has_emission = __float_as_int(d.z);
  return r;
}

__device__ __forceinline__ f3 fresnelSchlick(float cosTheta, f3 F0) {  // global.hpp:236-239
  float p = (float)pow(1.0 - (double)cosTheta, 5.0);
  return F0 + (mk(1.0f) - F0) * p;
}
__device__ __forceinline__ float fresnel(f3 Incident, f3 normal, float eta_i, float eta_t) {  // :242-260
  f3 I = normalized(Incident);
  f3 N = normalized(normal);
  float cosI_N = dot(I, N);
  if (cosI_N < 0) N = -N;
  float q = (eta_t - eta_i) / (eta_t + eta_i);
  float F0 = q * q;
  float m = 1 - dot(I, N);
  float m2 = m * m;
  return F0 + (1 - F0) * (m2 * m2 * m);
}
__device__ __forceinline__ f3 getReflectionDir(f3 incident, f3 normal) {  // :263-268
  f3 I = normalized(incident);
  f3 N = normalized(normal);
  return 2 * dot(N, I) * N - I;
}
__device__ __forceinline__ f3 getRefractionDir(f3 incident, f3 normal, float eta_i, float eta_t) {  // :271-301
  f3 I = normalized(incident);
  f3 N = normalized(normal);
  float cos_theta_i = dot(N, I);
  cos_theta_i = fmaxf(-1.f, fminf(1.f, cos_theta_i));
  if (cos_theta_i < 0) {
    N = -N;
    cos_theta_i = -cos_theta_i;
  }
  float sin_theta_i = sqrtf(1 - cos_theta_i * cos_theta_i);
  float sin_theta_t = (eta_i / eta_t) * sin_theta_i;
  if (sin_theta_i > (eta_t / eta_i)) return mk(0.f);
  float cos_theta_t = sqrtf(1 - sin_theta_t * sin_theta_t);
  return cos_theta_t * (-N) + eta_i / eta_t * (cos_theta_i * N - I);
}
__device__ __forceinline__ float D_ndf(f3 h, f3 n, float roughness) {  // :311-324
  float alpha = roughness * roughness;
  alpha = fmaxf(alpha, 1e-3f);
  float nh = dot(n, h);
  if (nh < 0) return 0;
  float cos_nh_2 = nh * nh;
  float sin_nh_2 = 1 - cos_nh_2;
  float sum = alpha * alpha * cos_nh_2 + sin_nh_2;
  if (sum == 0) return 1;
  return (alpha * alpha) / (T_PI * (sum * sum));
}
__device__ __forceinline__ float G1_term(float wh, float wn, float alpha) {
  float ang = acosf(wn);
  float tn = tanf(ang);
  return (((wh / wn) < 0) ? 0.f : 1.f) * 2.f / (1 + sqrtf(1 + alpha * alpha * (tn * tn)));
}
__device__ __forceinline__ float G_smf(f3 wi, f3 wo, f3 n, float roughness, f3 h) {  // :334-345
  float alpha = roughness * roughness;
  alpha = fmaxf(alpha, 1e-3f);
  float G1_wi = G1_term(dot(wi, h), dot(wi, n), alpha);
  float G1_wo = G1_term(dot(wo, h), dot(wo, n), alpha);
  if (isnan(G1_wi) || isnan(G1_wo)) return 0;
  return G1_wi * G1_wo;
}
__device__ __forceinline__ float getMisWeight(float pdf, float otherPdf) {  // :374-380
  return fdiv(pdf * pdf, (pdf + otherPdf) * (pdf + otherPdf));
}
__device__ __forceinline__ f3 SphereLocal2world(f3 n, f3 dir) {  // :387-409
  f3 N = normalized(n);
  f3 a = fabsf(N.x) > 0.9f ? mk(0.f, 1.f, 0.f) : mk(1.f, 0.f, 0.f);
  f3 S = normalized(cross(N, a));
  f3 T = cross(N, S);
  return normalized(dir.x * S + dir.y * T + dir.z * N);
}

// ---- Material.hpp ----------------------------------------------------------------------------
// (__noinline__ helpers take everything by value: a reference parameter would pin the caller's
// material / direction registers to local memory on the common Lambertian path as well)
static __device__ __noinline__ f3 BxDF_microfacet(const Mat m, f3 wi, f3 wo, f3 Ns, float eta_scene,
                                           bool TIR, float correctNormal) {
  if (m.type == TUTU_MAT_MICROFACET_R) {  // Material.hpp:87-108
    f3 h = normalized(wi + wo);
    float costheta = dot(h, wi);
    f3 F0 = mk(0.04f);
    F0 = mk(F0.x + m.metallic * (m.diffuse.x - F0.x), F0.y + m.metallic * (m.diffuse.y - F0.y),
            F0.z + m.metallic * (m.diffuse.z - F0.z));
    f3 F = fresnelSchlick(costheta, F0);
    float D = D_ndf(h, Ns, m.roughness);
    float G = G_smf(wi, wo, Ns, m.roughness, h);
    float denom = 4 * dot(wi, Ns) * dot(wo, Ns);
    if (denom == 0) return mk(0.f);
    f3 fr = (F * G * D) / denom;
    f3 diffuse_term = (mk(1.f) - F) * (m.diffuse / T_PI);
    return (diffuse_term + fr) * correctNormal;
  }
  // MICROFACET_T, Material.hpp:110-149
  float eta_i = eta_scene, eta_t = m.eta;
  f3 interN = Ns;
  if (dot(wo, Ns) < 0) {
    interN = -Ns;
    float s = eta_i;
    eta_i = eta_t;
    eta_t = s;
  }
  if (dot(wi, interN) >= 0) {
    f3 h = normalized(wo + wi);
    float F = fresnel(wi, h, eta_i, eta_t);
    if (TIR) F = 1.f;
    float D = D_ndf(h, interN, m.roughness);
    float G = G_smf(wi, wo, interN, m.roughness, h);
    float denom = 4 * dot(wi, interN) * dot(wo, interN);
    if (denom == 0) return mk(0.f);
    return mk((F * G * D) / denom) * correctNormal;
  }
  f3 h = -normalized(eta_i * wo + eta_t * wi);
  if (dot(h, interN) < 0) h = -h;
  float cos_ih = dot(wi, h), cos_oh = dot(wo, h), cos_in = dot(wi, interN), cos_on = dot(wo, interN);
  float F = fresnel(wi, h, eta_i, eta_t);
  float D = D_ndf(h, interN, m.roughness);
  float G = G_smf(wi, wo, interN, m.roughness, h);
  float numerator = fabsf(cos_ih) * fabsf(cos_oh) * eta_t * eta_t * (1 - F) * G * D;
  float s = eta_i * cos_ih + eta_t * cos_oh;
  float denominator = fabsf(cos_in) * fabsf(cos_on) * (s * s);
  if (denominator == 0) return mk(0.f);
  return mk(numerator / denominator * correctNormal);
}

static __device__ __noinline__ f3 BxDF_glass(const Mat m, f3 wi, f3 wo, f3 Ns, float eta_scene, bool TIR,
                                      float correctNormal) {  // Material.hpp:159-186
  f3 refDir = normalized(getReflectionDir(wo, Ns));
  float eta_i = eta_scene, eta_t = m.eta;
  f3 interN = Ns;
  if (dot(wo, Ns) < 0) {
    interN = -Ns;
    float s = eta_i;
    eta_i = eta_t;
    eta_t = s;
  }
  float F = fresnel(wi, interN, eta_i, eta_t);
  f3 transDir = normalized(getRefractionDir(wo, interN, eta_i, eta_t));
  interN = dot(interN, wi) < 0 ? -interN : interN;
  if (TIR) return mk(1 / dot(interN, wi) * correctNormal);
  if (FLOAT_EQUAL(dot(wi, refDir), 1.f)) return mk(F * 1 / dot(interN, wi) * correctNormal);
  if (FLOAT_EQUAL(dot(wi, transDir), 1.f)) return mk((1 - F) * 1 / dot(interN, wi) * correctNormal);
  return mk(0.f);
}

// Material::BxDF after its two-sidedness test and the adjoint swap (Material.hpp:74-191)
__device__ __forceinline__ f3 BxDF_core(const Mat& m, f3 wi, f3 wo, f3 Ng, f3 Ns, float eta_scene, bool TIR) {
  float correctNormal = fdiv(fabsf(dot(wi, Ns)), fabsf(dot(wi, Ng)));
  switch (m.type) {
    case TUTU_MAT_LAMBERTIAN: {
      float cos_theta = dot(wi, Ns);
      if (cos_theta >= 0.f) return m.diffuse * (T_INV_PI * correctNormal);
      return mk(0.f);
    }
    case TUTU_MAT_MICROFACET_R:
    case TUTU_MAT_MICROFACET_T:
      return BxDF_microfacet(m, wi, wo, Ns, eta_scene, TIR, correctNormal);
    case TUTU_MAT_PERFECT_REFLECTIVE: {
      if (FLOAT_EQUAL(dot(normalized(wi + wo), Ns), 1.f)) return mk(1 / fabsf(dot(Ns, wi)) * correctNormal);
      return mk(0.f);
    }
    case TUTU_MAT_PERFECT_REFRACTIVE:
      return BxDF_glass(m, wi, wo, Ns, eta_scene, TIR, correctNormal);
    default:
      return mk(0.f);
  }
}
__device__ __forceinline__ bool BxDF_sides_ok(const Mat& m, f3 wi, f3 wo, f3 Ng, f3 Ns) {  // Material.hpp:65-68
  if (m.type != TUTU_MAT_MICROFACET_T && m.type != TUTU_MAT_PERFECT_REFRACTIVE) {
    if (dot(wi, Ng) * dot(wi, Ns) <= 0 || dot(wo, Ng) * dot(wo, Ns) <= 0) return false;
  }
  return true;
}
__device__ __forceinline__ f3 BxDF(const Mat& m, f3 wi, f3 wo, f3 Ng, f3 Ns, float eta_scene,
                                   bool TIR = false) {  // Material.hpp:62-191 (adjoint = false)
  if (!BxDF_sides_ok(m, wi, wo, Ng, Ns)) return mk(0.f);
  return BxDF_core(m, wi, wo, Ng, Ns, eta_scene, TIR);
}
// adjoint = true (light sub-paths, BDPT.hpp:375,804,855): the side test sees the caller's wi/wo,
// everything after it the swapped pair (Material.hpp:70-73)
__device__ __forceinline__ f3 BxDF_adjoint(const Mat& m, f3 wi, f3 wo, f3 Ng, f3 Ns, float eta_scene,
                                           bool TIR = false) {
  if (!BxDF_sides_ok(m, wi, wo, Ng, Ns)) return mk(0.f);
  return BxDF_core(m, wo, wi, Ng, Ns, eta_scene, TIR);
}

// GGX half-vector in the local frame, shared by MICROFACET_R/T (Material.hpp:208-221, 232-242)
__device__ __forceinline__ f3 ggx_local_h(float r0, float r1, float a2) {
  float phi = 2 * T_PI * r1;
  float costheta = sqrtf((1 - r0) / (r0 * (a2 - 1) + 1));
  float sintheta = sqrtf(1 - costheta * costheta);
  float sp, cp;
  sincosf(phi, &sp, &cp);
  return normalized(mk(sintheta * cp, sintheta * sp, costheta));
}

struct DirSample {
  f3 wi;
  int flags;  // bit0 = success, bit1 = TIR ("special event")
};
// ra, rb, rc = getRandomFloat() calls in order.
static __device__ __noinline__ DirSample sampleDirection_special(const Mat m, f3 wo, f3 N, float eta_i,
                                                          float ra, float rb, float rc) {
  f3 out = mk(0.f);
  const int flags = [&]() -> int {
  switch (m.type) {
    case TUTU_MAT_MICROFACET_R: {  // Material.hpp:203-229
      if (dot(wo, N) <= 0.0f) return 0;
      float alhpa = m.roughness * m.roughness;
      float a2 = alhpa * fmaxf(m.alpha, 1e-3f);  // sic: opacity, Material.hpp:212-214
      f3 h = ggx_local_h(ra, rb, a2);
      f3 res = normalized(getReflectionDir(wo, SphereLocal2world(N, h)));
      if (dot(res, N) <= 0) return 0;
      out = res;
      return 1;
    }
    case TUTU_MAT_MICROFACET_T: {  // Material.hpp:231-268
      float a = fmaxf(m.roughness * m.roughness, 1e-3f);
      f3 h = ggx_local_h(ra, rb, a * a);
      float eta_t = m.eta;
      f3 interN = N;
      if (dot(wo, N) < 0) {
        float s = eta_i;
        eta_i = eta_t;
        eta_t = s;
        interN = -interN;
      }
      h = SphereLocal2world(interN, h);
      f3 res = getRefractionDir(wo, h, eta_i, eta_t);
      if (norm2(res) == 0) return 3;
      float F = fresnel(wo, h, eta_i, eta_t);
      out = (rc < F) ? getReflectionDir(wo, h) : res;
      return 1;
    }
    case TUTU_MAT_PERFECT_REFLECTIVE:  // Material.hpp:309-313
      out = getReflectionDir(wo, N);
      return 1;
    case TUTU_MAT_PERFECT_REFRACTIVE: {  // Material.hpp:314-336 (single draw)
      float eta_t = m.eta;
      f3 interN = N;
      if (dot(wo, N) < 0) {
        float s = eta_i;
        eta_i = eta_t;
        eta_t = s;
        interN = -interN;
      }
      f3 res = getRefractionDir(wo, interN, eta_i, eta_t);
      if (norm2(res) == 0) return 3;
      float F = fresnel(wo, interN, eta_i, eta_t);
      out = (ra < F) ? getReflectionDir(wo, interN) : res;
      return 1;
    }
    default:
      return 0;
  }
  }();
  return DirSample{out, flags};
}

__device__ __forceinline__ int sampleDirection(const Mat& m, f3 wo, f3 N, f3& out, float eta_i, float ra,
                                               float rb, float rc) {
  if (m.type == TUTU_MAT_LAMBERTIAN) {  // Material.hpp:270-307
    if (dot(wo, N) <= 0.0f) return 0;
    float cosTheta = sqrtf(ra);
    float phi = 2 * T_PI * rb;
    float sinTheta = sqrtf(fmaxf(0.f, 1.f - ra));
    float sp, cp;
    sincosf(phi, &sp, &cp);
    f3 dir = normalized(mk(cp * sinTheta, sp * sinTheta, cosTheta));
    f3 res = SphereLocal2world(N, dir);
    if (dot(normalized(res), N) < 0) return 0;
    out = res;
    return 1;
  }
  const DirSample ds = sampleDirection_special(m, wo, N, eta_i, ra, rb, rc);
  if (ds.flags & 1) out = ds.wi;
  return ds.flags;
}

static __device__ __noinline__ float pdf_special(const Mat m, f3 wi, f3 wo, f3 N, float eta_i, float eta_t) {
  switch (m.type) {
    case TUTU_MAT_MICROFACET_R: {  // Material.hpp:362-373
      f3 h = normalized(wo + wi);
      float cosTheta = fmaxf(dot(N, h), 0.f);
      return D_ndf(h, N, m.roughness) * cosTheta / (4.f * dot(wo, h));
    }
    case TUTU_MAT_MICROFACET_T: {  // Material.hpp:374-405
      f3 interN = N;
      if (dot(wo, N) < 0) {
        interN = -N;
        float s = eta_i;
        eta_i = eta_t;
        eta_t = s;
      }
      float F = fresnel(wo, interN, eta_i, eta_t);
      if (dot(wi, interN) >= 0) {
        f3 h = normalized(wo + wi);
        float cosTheta = fabsf(dot(interN, h));
        float deno = 4.f * dot(wo, h);
        if (deno == 0) return 0;
        return F * D_ndf(h, interN, m.roughness) * cosTheta / deno;
      }
      f3 h = -normalized(eta_i * wo + eta_t * wi);
      float cosTheta = dot(interN, h);
      if (cosTheta < 0) {
        h = -h;
        cosTheta = fabsf(cosTheta);
      }
      float ds = eta_i * dot(wi, h) + eta_t * dot(wo, h);
      float jacobian = (eta_t * eta_t * fabsf(dot(wo, h))) / (ds * ds);
      if (ds == 0) return 0;
      return (1 - F) * D_ndf(h, interN, m.roughness) * cosTheta * jacobian;
    }
    case TUTU_MAT_PERFECT_REFLECTIVE:  // Material.hpp:407-412
      return FLOAT_EQUAL(dot(normalized(wi + wo), N), 1.f) ? 1.f : 0.f;
    case TUTU_MAT_PERFECT_REFRACTIVE: {  // Material.hpp:414-432
      f3 refDir = normalized(getReflectionDir(wo, N));
      f3 nDir = N;
      if (dot(wo, nDir) < 0) {
        float s = eta_i;
        eta_i = eta_t;
        eta_t = s;
        nDir = -N;
      }
      f3 transDir = normalized(getRefractionDir(wo, nDir, eta_i, eta_t));
      float F = fresnel(wo, nDir, eta_i, eta_t);
      if (FLOAT_EQUAL(dot(wi, refDir), 1.f)) return F;
      if (FLOAT_EQUAL(dot(wi, transDir), 1.f)) return 1 - F;
      return 0;
    }
    default:
      return 1;
  }
}
__device__ __forceinline__ float mat_pdf_eval(const Mat& m, f3 wi, f3 wo, f3 N, float eta_i, float eta_t) {
  if (m.type == TUTU_MAT_LAMBERTIAN) {  // Material.hpp:353-360
    float c = dot(wi, N);
    return c > 0.0f ? c * T_INV_PI : 0.0f;
  }
  return pdf_special(m, wi, wo, N, eta_i, eta_t);
}

// ---- textures --------------------------------------------------------------------------------
__device__ __forceinline__ f3 tex_fetch(const DevScene& sc, int channel, int index, float u, float v) {  // Texture.hpp:18-39
  const int4 h = __ldg(sc.tex_headers[channel] + index);
  if ((h.x == 0 && h.y == 0) || h.w <= 0) return mk(0.f);
  if (u > 0) u = u - (int)u;
  else
    u = 1 - (fabsf(u) - (int)fabsf(u));
  if (v > 0) v = v - (int)v;
  else
    v = 1 - (fabsf(v) - (int)fabsf(v));
  int x = (int)(u * h.x);
  int y = (int)(v * h.y);
  int idx = y * h.x + x;
  if (idx < 0) idx = 0;
  if (idx >= h.w) idx = h.w - 1;
  const float4 t = __ldg(sc.texels + (size_t)h.z + idx);
  return mk(t.x, t.y, t.z);
}

}  // namespace tutu
