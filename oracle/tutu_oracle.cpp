// tutu_oracle — CPU restatement of the reference's hot path (bobhansky/TutuRenderer).
//
// TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg as the checker.  The product library never links or calls this file.
//
// Parity pin: the reference ships no tests or golden vectors, so this port is
// pinned against the reference itself, compiled here as oracle/_ref/ref_harness
// (oracle/Makefile): tree topology and ray-batch hits must be bit-identical, renders must agree
// statistically (tests/tools/make_golden.py wrote tests/golden/*, tests/test_oracle_vs_reference.py
// re-checks live when oracle/_ref exists).
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// The traversal is the reference's literal recursion (both children, no t-pruning) and the path
// tracer is the literal recursive traceRay; only the random numbers differ: the reference draws
// from a thread-local mt19937 (global.hpp:182-199), here each call site reads a fixed slot of a
// Philox4x32-10 stream keyed by (seed; pixel, sample, depth), the same stream the CUDA path uses,
// so GPU and oracle trace the same paths.
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off (no FMA, like the reference build).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <tuple>
#include <vector>

#include "tutu_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Vector.hpp:70-225
// ---------------------------------------------------------------------------------------------
struct V2 {
  float x = 0, y = 0;
};
struct V3 {
  float x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(float a) : x(a), y(a), z(a) {}
  V3(float a, float b, float c) : x(a), y(b), z(c) {}
  V3 operator*(float c) const { return V3(x * c, y * c, z * c); }
  V3 operator/(float c) const { return V3(x / c, y / c, z / c); }
  V3 operator*(const V3& v) const { return V3(x * v.x, y * v.y, z * v.z); }
  V3 operator/(const V3& v) const { return V3(x / v.x, y / v.y, z / v.z); }
  V3 operator-(const V3& v) const { return V3(x - v.x, y - v.y, z - v.z); }
  V3 operator+(const V3& v) const { return V3(x + v.x, y + v.y, z + v.z); }
  V3 operator-() const { return V3(-x, -y, -z); }
  float dot(const V3& v) const { return x * v.x + y * v.y + z * v.z; }  // :186
  float norm() const { return sqrtf(x * x + y * y + z * z); }
  float norm2() const { return x * x + y * y + z * z; }
};
inline V3 operator*(float c, const V3& v) { return V3(v.x * c, v.y * c, v.z * c); }
inline V3 operator-(float c, const V3& v) { return V3(c - v.x, c - v.y, c - v.z); }
inline V3 normalized(const V3& v) {  // :213-220
  float mag = sqrtf((v.x * v.x + v.y * v.y + v.z * v.z));
  if (mag > 0) {
    float mag_inv = 1 / mag;
    return V3(v.x * mag_inv, v.y * mag_inv, v.z * mag_inv);
  }
  return v;
}
inline V3 crossProduct(const V3& a, const V3& b) {  // :223
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// ---------------------------------------------------------------------------------------------
// global.hpp
// ---------------------------------------------------------------------------------------------
#define O_PI 3.1415926535897f   // :15
#define O_EPSILON 0.0005f       // :16
#define O_MIN_DIVISOR 0.04f     // :26
#define O_MAX_DEPTH 6           // PathTracing.hpp:5
#define O_MIN_DEPTH 3           // PathTracing.hpp:6

inline bool FLOAT_EQUAL(float x, float y) { return fabsf(x - y) < 0.0001f; }  // :133-135
inline float clampf(float lo, float hi, float v) { return std::max(lo, std::min(hi, v)); }  // :50
inline V3 lerp3(const V3& v0, const V3& v1, float x) {  // :42-48
  return V3(v0.x + x * (v1.x - v0.x), v0.y + x * (v1.y - v0.y), v0.z + x * (v1.z - v0.z));
}

void solveQuadratic(float& t1, float& t2, float A, float B, float C) {  // :147-167
  float discriminant = B * B - 4 * A * C;
  if (discriminant < 0) {
    t1 = FLT_MAX;
    t2 = FLT_MAX;
  } else if (discriminant == 0) {
    t1 = (-B + sqrtf(discriminant)) / (2 * A);
    t2 = t1;
  } else {
    t1 = (-B + sqrtf(discriminant)) / (2 * A);
    t2 = (-B - sqrtf(discriminant)) / (2 * A);
  }
  if (t1 > t2) std::swap(t1, t2);
}

V3 fresnelSchlick(float cosTheta, const V3& F0) {  // :236-239 (double pow, rounded to float)
  float p = (float)pow(1.0 - (double)cosTheta, 5.0);
  return F0 + (1.0f - F0) * p;
}

float fresnel(const V3& Incident, const V3& normal, float eta_i, float eta_t) {  // :242-260
  V3 I = normalized(Incident);
  V3 N = normalized(normal);
  float cosI_N = I.dot(N);
  if (cosI_N < 0) N = -N;
  float F0 = powf(((eta_t - eta_i) / (eta_t + eta_i)), 2.f);
  float Fr = F0 + (1 - F0) * (powf(1 - (I.dot(N)), 5.f));
  return Fr;
}

V3 getReflectionDir(const V3& incident, const V3& normal) {  // :263-268
  V3 I = normalized(incident);
  V3 N = normalized(normal);
  return 2 * (N.dot(I)) * N - I;
}

V3 getRefractionDir(const V3& incident, const V3& normal, float eta_i, float eta_t) {  // :271-301
  V3 I = normalized(incident);
  V3 N = normalized(normal);
  float cos_theta_i = N.dot(I);
  cos_theta_i = clampf(-1, 1, cos_theta_i);
  if (cos_theta_i < 0) {
    N = -N;
    cos_theta_i = -cos_theta_i;
  }
  float sin_theta_i = sqrtf(1 - powf(cos_theta_i, 2));
  float sin_theta_t = (eta_i / eta_t) * sin_theta_i;
  if (sin_theta_i > (eta_t / eta_i)) return V3(0);
  float cos_theta_t = sqrtf(1 - powf(sin_theta_t, 2));
  return cos_theta_t * (-N) + eta_i / eta_t * (cos_theta_i * N - I);
}

float D_ndf(const V3& h, const V3& n, float roughness) {  // :311-324
  float alpha = roughness * roughness;
  alpha = std::max(alpha, 1e-3f);
  if (n.dot(h) < 0) return 0;
  float cos_nh_2 = (n.dot(h)) * (n.dot(h));
  float sin_nh_2 = 1 - cos_nh_2;
  float sum = alpha * alpha * cos_nh_2 + sin_nh_2;
  if (sum == 0) return 1;
  float res = (alpha * alpha) / (O_PI * (sum * sum));
  return res;
}

float G_smf(const V3& wi, const V3& wo, const V3& n, float roughness, const V3& h) {  // :334-345
  float alpha = roughness * roughness;
  alpha = std::max(alpha, 1e-3f);
  float angle_wi_n = acosf(wi.dot(n));
  float angle_wo_n = acosf(wo.dot(n));
  float G1_wi = ((wi.dot(h) / wi.dot(n)) < 0 ? 0 : 1) * 2 /
                (1 + sqrtf(1 + alpha * alpha * powf(tanf(angle_wi_n), 2)));
  float G1_wo = ((wo.dot(h) / wo.dot(n)) < 0 ? 0 : 1) * 2 /
                (1 + sqrtf(1 + alpha * alpha * powf(tanf(angle_wo_n), 2)));
  if (std::isnan(G1_wi) || std::isnan(G1_wo)) return 0;
  return G1_wi * G1_wo;
}

float getMisWeight(float pdf, float otherPdf) {  // :374-380
  return (pdf * pdf) / ((pdf + otherPdf) * (pdf + otherPdf));
}

void offsetRayOrig(V3& orig, V3 interNormal, bool rayIsInside = false) {  // :383-385
  if (rayIsInside)
    orig = orig - interNormal * O_EPSILON;
  else
    orig = orig + interNormal * O_EPSILON;
}

V3 SphereLocal2world(const V3& n, const V3& dir) {  // :387-409
  V3 a;
  V3 N = normalized(n);
  if (fabsf(N.x) > 0.9f)
    a = V3(0.f, 1.f, 0.f);
  else
    a = V3(1.f, 0.f, 0.f);
  V3 S = normalized(crossProduct(N, a));
  V3 T = crossProduct(N, S);
  return normalized(dir.x * S + dir.y * T + dir.z * N);
}

// ---------------------------------------------------------------------------------------------
// random numbers: Philox4x32-10, counter (pixel, sample, depth, block), key = seed
// slots per depth: 0 light index, 1-2 light point, 3-4 BSDF, 5 third BSDF draw / roulette
// Path tracing (packed6): the six numbers of a vertex are the six 21-bit fields of ONE block's 128 bits
// (u = field * 2^-21), the GPU's draw6; BDPT keeps one 24-bit number per 32-bit word (draw4).
// ---------------------------------------------------------------------------------------------
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
struct Rng {
  uint64_t seed;
  uint32_t pixel, sample;
  bool packed6 = false;
  float get(int depth, int slot) const {
    if (packed6) {
      uint32_t c[4] = {pixel, sample, (uint32_t)depth, 0u};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      const uint32_t hi = c[slot < 3 ? 0 : 2], lo = c[slot < 3 ? 1 : 3];
      uint32_t f;
      switch (slot % 3) {
        case 0: f = hi >> 11; break;
        case 1: f = ((hi & 0x7FFu) << 10) | (lo >> 22); break;
        default: f = (lo >> 1) & 0x1FFFFFu; break;
      }
      return (float)f * (1.0f / 2097152.0f);
    }
    uint32_t c[4] = {pixel, sample, (uint32_t)depth, (uint32_t)(slot >> 2)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (float)(c[slot & 3] >> 8) * (1.0f / 16777216.0f);
  }
};

// ---------------------------------------------------------------------------------------------
// Material.hpp
// ---------------------------------------------------------------------------------------------
struct Material {
  V3 diffuse = V3(0.9f, 0.9f, 0.9f);
  V3 specular = V3(1.f);
  V3 emission = V3(0.f);
  int mType = TUTU_MAT_LAMBERTIAN;
  float alpha = 1, eta = 1, roughness = 1, metallic = 0;

  bool hasEmission() const { return emission.x || emission.y || emission.z; }  // :54-56

  // :62-191
  V3 BxDF(const V3& wi_, const V3& wo_, const V3& Ng, const V3& Ns, float eta_scene,
          bool adjoint = false, bool TIR = false) const {
    V3 wi = wi_;
    V3 wo = wo_;
    if (mType != TUTU_MAT_MICROFACET_T && mType != TUTU_MAT_PERFECT_REFRACTIVE) {
      if (wi.dot(Ng) * wi.dot(Ns) <= 0 || wo.dot(Ng) * wo.dot(Ns) <= 0) return 0;
    }
    if (adjoint) {
      wi = wo_;
      wo = wi_;
    }
    float correctNormal = fabsf(wi.dot(Ns)) / fabsf(wi.dot(Ng));
    switch (mType) {
      case TUTU_MAT_LAMBERTIAN: {
        float cos_theta = wi.dot(Ns);
        if (cos_theta >= 0.f) return diffuse / O_PI * correctNormal;
        return V3(0.f);
      }
      case TUTU_MAT_MICROFACET_R: {
        V3 h = normalized(wi + wo);
        float costheta = h.dot(wi);
        V3 F0(0.04f);
        F0 = lerp3(F0, diffuse, metallic);
        V3 F = fresnelSchlick(costheta, F0);
        float D = D_ndf(h, Ns, roughness);
        float G = G_smf(wi, wo, Ns, roughness, h);
        float denom = 4 * wi.dot(Ns) * wo.dot(Ns);
        if (denom == 0) return 0;
        V3 fr = (F * G * D) / denom;
        V3 diffuse_term = (1.f - F) * (diffuse / O_PI);
        return (diffuse_term + fr) * correctNormal;
      }
      case TUTU_MAT_MICROFACET_T: {
        float eta_i = eta_scene;
        float eta_t = eta;
        V3 interN = Ns;
        if (wo.dot(Ns) < 0) {
          interN = -Ns;
          std::swap(eta_i, eta_t);
        }
        if (wi.dot(interN) >= 0) {
          V3 h = normalized(wo + wi);
          float F = fresnel(wi, h, eta_i, eta_t);
          if (TIR) F = 1.f;
          float D = D_ndf(h, interN, roughness);
          float G = G_smf(wi, wo, interN, roughness, h);
          float denom = 4 * wi.dot(interN) * wo.dot(interN);
          if (denom == 0) return 0;
          V3 fr = V3((F * G * D) / denom);
          return fr * correctNormal;
        } else {
          V3 h = -normalized(eta_i * wo + eta_t * wi);
          if (h.dot(interN) < 0) h = -h;
          float cos_ih = wi.dot(h), cos_oh = wo.dot(h), cos_in = wi.dot(interN),
                cos_on = wo.dot(interN);
          float F = fresnel(wi, h, eta_i, eta_t);
          float D = D_ndf(h, interN, roughness);
          float G = G_smf(wi, wo, interN, roughness, h);
          float numerator = fabsf(cos_ih) * fabsf(cos_oh) * eta_t * eta_t * (1 - F) * G * D;
          float denominator =
              fabsf(cos_in) * fabsf(cos_on) * powf(eta_i * cos_ih + eta_t * cos_oh, 2);
          if (denominator == 0) return 0;
          return V3(numerator / denominator * correctNormal);
        }
      }
      case TUTU_MAT_PERFECT_REFLECTIVE: {
        if (FLOAT_EQUAL(normalized(wi + wo).dot(Ns), 1.f))
          return V3(1 / fabsf(Ns.dot(wi)) * correctNormal);
        return 0;
      }
      case TUTU_MAT_PERFECT_REFRACTIVE: {
        V3 refDir = normalized(getReflectionDir(wo, Ns));
        float eta_i = eta_scene;
        float eta_t = this->eta;
        float F;
        V3 interN = Ns;
        if (wo.dot(Ns) < 0) {
          interN = -Ns;
          std::swap(eta_i, eta_t);
        }
        F = fresnel(wi, interN, eta_i, eta_t);
        V3 transDir = normalized(getRefractionDir(wo, interN, eta_i, eta_t));
        interN = interN.dot(wi) < 0 ? -interN : interN;
        if (TIR) return V3(1 / interN.dot(wi) * correctNormal);
        if (FLOAT_EQUAL(wi.dot(refDir), 1.f)) return V3(F * 1 / interN.dot(wi) * correctNormal);
        else if (FLOAT_EQUAL(wi.dot(transDir), 1.f))
          return V3((1 - F) * 1 / interN.dot(wi) * correctNormal);
        return V3(0.f);
      }
      default:
        return V3(0.f);
    }
  }

  // :200-343.  r(k) = k-th getRandomFloat() call of this invocation.
  template <class R>
  std::tuple<bool, bool> sampleDirection(const V3& wo, const V3& N, V3& sampledRes, float eta_i,
                                         R&& r) {
    switch (mType) {
      case TUTU_MAT_MICROFACET_R: {
        if (wo.dot(N) <= 0.0f) return {false, false};
        float r0 = r(0);
        float r1 = r(1);
        float alhpa = roughness * roughness;
        alpha = std::max(alpha, 1e-3f);  // sic (:212-214): clamps the opacity member
        float a2 = alhpa * alpha;
        float phi = 2 * O_PI * r1;
        float costheta = sqrtf((1 - r0) / (r0 * (a2 - 1) + 1));
        float sintheta = sqrtf(1 - costheta * costheta);
        float rr = sintheta;
        V3 h = normalized(V3(rr * cosf(phi), rr * sinf(phi), costheta));
        V3 res = getReflectionDir(wo, SphereLocal2world(N, h));
        res = normalized(res);
        if (res.dot(N) <= 0) return {false, false};
        sampledRes = res;
        return {true, false};
      }
      case TUTU_MAT_MICROFACET_T: {
        float r0 = r(0);
        float r1 = r(1);
        float a = roughness * roughness;
        a = std::max(a, 1e-3f);
        float a2 = a * a;
        float phi = 2 * O_PI * r1;
        float costheta = sqrtf((1 - r0) / (r0 * (a2 - 1) + 1));
        float sintheta = sqrtf(1 - costheta * costheta);
        float rr = sintheta;
        V3 h = normalized(V3(rr * cosf(phi), rr * sinf(phi), costheta));
        float eta_t = eta;
        V3 interN = N;
        if (wo.dot(N) < 0) {
          std::swap(eta_i, eta_t);
          interN = -interN;
        }
        h = SphereLocal2world(interN, h);
        V3 res = getRefractionDir(wo, h, eta_i, eta_t);
        if (res.norm2() == 0) return {true, true};
        float F = fresnel(wo, h, eta_i, eta_t);
        if (r(2) < F)
          sampledRes = getReflectionDir(wo, h);
        else
          sampledRes = res;
        return {true, false};
      }
      case TUTU_MAT_LAMBERTIAN: {
        if (wo.dot(N) <= 0.0f) return {false, false};
        float r1 = r(0);
        float r2 = r(1);
        float cosTheta = sqrtf(r1);
        float phi = 2 * O_PI * r2;
        V3 dir;
        float sinTheta = sqrtf(std::max(0.f, 1.f - r1));
        dir.x = cosf(phi) * sinTheta;
        dir.y = sinf(phi) * sinTheta;
        dir.z = cosTheta;
        dir = normalized(dir);
        V3 res = SphereLocal2world(N, dir);
        if (normalized(res).dot(N) < 0) return {false, false};
        sampledRes = res;
        return {true, false};
      }
      case TUTU_MAT_PERFECT_REFLECTIVE: {
        sampledRes = getReflectionDir(wo, N);
        return {true, false};
      }
      case TUTU_MAT_PERFECT_REFRACTIVE: {
        float eta_t = eta;
        V3 interN = N;
        if (wo.dot(N) < 0) {
          std::swap(eta_i, eta_t);
          interN = -interN;
        }
        V3 res = getRefractionDir(wo, interN, eta_i, eta_t);
        if (res.norm2() == 0) return {true, true};
        float F = fresnel(wo, interN, eta_i, eta_t);
        if (r(0) < F)
          sampledRes = getReflectionDir(wo, interN);
        else
          sampledRes = res;
        return {true, false};
      }
      default:
        return {false, false};
    }
  }

  // :350-439
  float pdf(const V3& wi, const V3& wo, const V3& N, float eta_i = 1.f, float eta_t = 1.f) const {
    switch (mType) {
      case TUTU_MAT_LAMBERTIAN: {
        if (wi.dot(N) > 0.0f) return wi.dot(N) / O_PI;
        return 0.0f;
      }
      case TUTU_MAT_MICROFACET_R: {
        V3 h = normalized(wo + wi);
        float cosTheta = N.dot(h);
        cosTheta = std::max(cosTheta, 0.f);
        return D_ndf(h, N, roughness) * cosTheta / (4.f * wo.dot(h));
      }
      case TUTU_MAT_MICROFACET_T: {
        V3 interN = N;
        if (wo.dot(N) < 0) {
          interN = -N;
          std::swap(eta_i, eta_t);
        }
        float F = fresnel(wo, interN, eta_i, eta_t);
        if (wi.dot(interN) >= 0) {
          V3 h = normalized(wo + wi);
          float cosTheta = interN.dot(h);
          cosTheta = fabsf(cosTheta);
          float deno = 4.f * wo.dot(h);
          if (deno == 0) return 0;
          return F * D_ndf(h, interN, roughness) * cosTheta / deno;
        } else {
          V3 h = -normalized(eta_i * wo + eta_t * wi);
          float cosTheta = interN.dot(h);
          if (cosTheta < 0) {
            h = -h;
            cosTheta = fabsf(cosTheta);
          }
          float denominatorSqrt = eta_i * wi.dot(h) + eta_t * wo.dot(h);
          float jacobian = (eta_t * eta_t * fabsf(wo.dot(h))) / (denominatorSqrt * denominatorSqrt);
          if (denominatorSqrt == 0) return 0;
          return (1 - F) * D_ndf(h, interN, roughness) * cosTheta * jacobian;
        }
      }
      case TUTU_MAT_PERFECT_REFLECTIVE: {
        if (FLOAT_EQUAL(normalized(wi + wo).dot(N), 1.f)) return 1;
        return 0;
      }
      case TUTU_MAT_PERFECT_REFRACTIVE: {
        V3 refDir = normalized(getReflectionDir(wo, N));
        V3 nDir = N;
        if (wo.dot(nDir) < 0) {
          std::swap(eta_i, eta_t);
          nDir = -N;
        }
        V3 transDir = normalized(getRefractionDir(wo, nDir, eta_i, eta_t));
        float F = fresnel(wo, nDir, eta_i, eta_t);
        if (FLOAT_EQUAL(wi.dot(refDir), 1.f)) return F;
        else if (FLOAT_EQUAL(wi.dot(transDir), 1.f))
          return 1 - F;
        return 0;
      }
      default:
        return 1;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// BoundBox.hpp
// ---------------------------------------------------------------------------------------------
struct BoundBox {
  V3 pMin, pMax;
  BoundBox() {}
  BoundBox(const V3& p1, const V3& p2) {  // :13-27
    pMin = V3(fminf(p1.x, p2.x), fminf(p1.y, p2.y), fminf(p1.z, p2.z));
    pMax = V3(fmaxf(p1.x, p2.x), fmaxf(p1.y, p2.y), fmaxf(p1.z, p2.z));
  }
  V3 Centroid() const { return 0.5f * pMin + 0.5f * pMax; }  // :35
  int maxExtent() const {                                     // :43-52
    V3 d = pMax - pMin;
    if (d.x > d.y && d.x > d.z) return 0;
    else if (d.y > d.z)
      return 1;
    else
      return 2;
  }
  bool IntersectRay(const V3& rayOrig, const V3& rayDir) const {  // :55-92
    V3 invDir = {1 / rayDir.x, 1 / rayDir.y, 1 / rayDir.z};
    float tmin_x = (pMin.x - rayOrig.x) * invDir.x;
    float tmax_x = (pMax.x - rayOrig.x) * invDir.x;
    float tmin_y = (pMin.y - rayOrig.y) * invDir.y;
    float tmax_y = (pMax.y - rayOrig.y) * invDir.y;
    float tmin_z = (pMin.z - rayOrig.z) * invDir.z;
    float tmax_z = (pMax.z - rayOrig.z) * invDir.z;
    if (rayDir.x < 0) std::swap(tmin_x, tmax_x);
    if (rayDir.y < 0) std::swap(tmin_y, tmax_y);
    if (rayDir.z < 0) std::swap(tmin_z, tmax_z);
    float t_enter, t_exit;
    float buffer = tmin_y > tmin_z ? tmin_y : tmin_z;
    t_enter = tmin_x > buffer ? tmin_x : buffer;
    buffer = tmax_y < tmax_z ? tmax_y : tmax_z;
    t_exit = tmax_x < buffer ? tmax_x : buffer;
    if (t_enter <= t_exit && t_exit >= 0.f) return true;
    return false;
  }
};
BoundBox Union(const BoundBox& b1, const BoundBox& b2) {  // :97-109
  V3 mn(fminf(b1.pMin.x, b2.pMin.x), fminf(b1.pMin.y, b2.pMin.y), fminf(b1.pMin.z, b2.pMin.z));
  V3 mx(fmaxf(b1.pMax.x, b2.pMax.x), fmaxf(b1.pMax.y, b2.pMax.y), fmaxf(b1.pMax.z, b2.pMax.z));
  return BoundBox(mn, mx);
}
BoundBox Union(const BoundBox& b, const V3& v) {  // :112-124
  V3 mn(fminf(b.pMin.x, v.x), fminf(b.pMin.y, v.y), fminf(b.pMin.z, v.z));
  V3 mx(fmaxf(b.pMax.x, v.x), fmaxf(b.pMax.y, v.y), fmaxf(b.pMax.z, v.z));
  return BoundBox(mn, mx);
}

// ---------------------------------------------------------------------------------------------
// Intersection.hpp, Object.hpp, Triangle.hpp, Sphere.hpp
// ---------------------------------------------------------------------------------------------
struct Object;
struct Intersection {
  bool intersected = false;
  float t = FLT_MAX;
  V3 pos, Ng, Ns;
  V2 textPos;
  int diffuseIndex = -1, normalMapIndex = -1, roughnessMapIndex = -1, metallicMapIndex = -1;
  Material mtlcolor;
  const Object* obj = nullptr;
  float u = 0, v = 0;  // not in the reference record: locals of Triangle::intersect, kept for parity
};

struct Object {
  int objectType = TUTU_PRIM_TRIANGLE;
  int index = -1;  // Scene::objList position
  Material mtlcolor;
  bool isTextureActivated = false;
  int textureIndex = -1, normalMapIndex = -1, roughnessMapIndex = -1, metallicMapIndex = -1;
  BoundBox bound;
  // triangle
  V3 v0, v1, v2, n0, n1, n2;
  V2 uv0, uv1, uv2;
  // sphere
  V3 centerPos;
  float radius = 1.f;

  void initializeBound() {
    if (objectType == TUTU_PRIM_SPHERE) {  // Sphere.hpp:129-133
      V3 mn = {centerPos.x - radius, centerPos.y - radius, centerPos.z - radius};
      V3 mx = {centerPos.x + radius, centerPos.y + radius, centerPos.z + radius};
      bound = BoundBox(mn, mx);
    } else {  // Triangle.hpp:104-107
      bound = BoundBox(v0, v1);
      bound = Union(bound, v2);
    }
  }

  float getArea() const {
    if (objectType == TUTU_PRIM_SPHERE) return radius * radius * O_PI;  // Sphere.hpp:135-137 (sic)
    V3 e1 = v1 - v0;                                                    // Triangle.hpp:109-116
    V3 e2 = v2 - v0;
    return crossProduct(e1, e2).norm() * 0.5f;
  }

  bool intersect(const V3& orig, const V3& dir, Intersection& inter) const {
    return objectType == TUTU_PRIM_SPHERE ? intersectSphere(orig, dir, inter)
                                          : intersectTriangle(orig, dir, inter);
  }

  bool intersectTriangle(const V3& orig, const V3& dir, Intersection& inter) const {  // Triangle.hpp:23-74
    V3 E1 = v1 - v0;
    V3 E2 = v2 - v0;
    V3 S = orig - v0;
    V3 S1 = crossProduct(dir, E2);
    V3 S2 = crossProduct(S, E1);
    V3 normal = crossProduct(E1, E2);
    normal = normalized(normal);
    if (FLOAT_EQUAL(dir.dot(normal), 0.f)) return false;
    V3 rightVec(S2.dot(E2), S1.dot(S), S2.dot(dir));
    if (S1.dot(E1) == 0.f) return false;
    float left = 1.0f / S1.dot(E1);
    V3 res = left * rightVec;
    if (res.x > 0 && 1 - res.y - res.z > 0 && res.y > 0 && res.z > 0) {
      inter.intersected = true;
      inter.obj = this;
      inter.t = res.x;
      inter.pos = orig + inter.t * dir;
      inter.mtlcolor = this->mtlcolor;
      inter.Ns = normalized((n0 * (1 - res.y - res.z)) + n1 * res.y + n2 * res.z);
      inter.Ng = normal;
      inter.u = res.y;
      inter.v = res.z;
      if (isTextureActivated) {
        float w = 1 - res.y - res.z;
        inter.textPos.x = uv0.x * w + uv1.x * res.y + uv2.x * res.z;  // Vector2f ops, Vector.hpp:57-65
        inter.textPos.y = uv0.y * w + uv1.y * res.y + uv2.y * res.z;
        inter.diffuseIndex = this->textureIndex;
        inter.normalMapIndex = this->normalMapIndex;
        inter.roughnessMapIndex = roughnessMapIndex;
        inter.metallicMapIndex = metallicMapIndex;
      }
      return true;
    }
    return false;
  }

  void fillSphere(const V3& orig, const V3& dir, Intersection& inter) const {  // Sphere.hpp:50-80,96-122
    inter.intersected = true;
    inter.obj = this;
    inter.mtlcolor = this->mtlcolor;
    inter.pos = orig + inter.t * dir;
    inter.Ng = normalized(inter.pos - centerPos);
    inter.Ns = inter.Ng;
    inter.u = inter.v = 0;
    if (isTextureActivated) {
      float u, v;
      float phi = acosf(inter.Ng.z);
      v = phi / O_PI;
      float theta = atan2f(inter.Ng.y, inter.Ng.x);
      if (theta < 0) theta += 2 * O_PI;
      u = (theta / (2.f * O_PI));
      inter.textPos.x = u;
      inter.textPos.y = v;
      inter.diffuseIndex = this->textureIndex;
      inter.normalMapIndex = normalMapIndex;
      inter.roughnessMapIndex = roughnessMapIndex;
      inter.metallicMapIndex = metallicMapIndex;
    }
  }

  bool intersectSphere(const V3& orig, const V3& dir, Intersection& inter) const {  // Sphere.hpp:26-126
    float A = 1.f;
    float B = 2 * (dir.x * (orig.x - centerPos.x) + dir.y * (orig.y - centerPos.y) +
                   dir.z * (orig.z - centerPos.z));
    // pow(float,int) is evaluated in double (C++11 promotion); the sum is rounded once
    float C = (float)(pow((double)(orig.x - centerPos.x), 2) + pow((double)(orig.y - centerPos.y), 2) +
                      pow((double)(orig.z - centerPos.z), 2) - (double)(radius * radius));
    float t1 = 0;
    float t2 = 0;
    solveQuadratic(t1, t2, A, B, C);
    inter.intersected = false;
    if (FLOAT_EQUAL(t1, FLT_MAX) && FLOAT_EQUAL(t2, FLT_MAX)) {
      return false;
    } else if (FLOAT_EQUAL(t1, t2)) {
      if (t1 < 0) return false;
      inter.t = t1;
      fillSphere(orig, dir, inter);
      return true;
    } else {
      if (t1 > 0 && t2 > 0) inter.t = t1;
      else if (t1 > 0 && t2 < 0)
        inter.t = t1;
      else if (t1 < 0 && t2 > 0)
        inter.t = t2;
      else
        return false;
      fillSphere(orig, dir, inter);
      return true;
    }
  }

  // Triangle.hpp:119-142, Sphere.hpp:139-164; ra, rb = the two getRandomFloat() draws in order
  void samplePoint(Intersection& inter, float& pdf, float ra, float rb) const {
    if (objectType == TUTU_PRIM_SPHERE) {
      float theta = ra * 2 * O_PI;
      float phi = rb * O_PI;
      inter.pos.x = centerPos.x + radius * cosf(theta) * sinf(phi);
      inter.pos.y = centerPos.y + radius * sinf(theta) * sinf(phi);
      inter.pos.z = centerPos.z + radius * cosf(phi);
      inter.Ng = normalized(inter.pos - centerPos);
      inter.Ns = inter.Ng;
      inter.intersected = true;
      inter.mtlcolor = mtlcolor;
      inter.obj = this;
      pdf = 1 / getArea();
      return;
    }
    float u = ra;
    float v = rb * (1 - u);
    V3 pos = (1 - u - v) * v0 + u * v1 + v * v2;
    inter.pos = pos;
    inter.Ng = (1 - u - v) * n0 + u * n1 + v * n2;
    inter.Ng = normalized(inter.Ng);
    inter.Ns = inter.Ng;
    inter.intersected = true;
    inter.mtlcolor = mtlcolor;
    inter.obj = this;
    float area = getArea();
    pdf = 1.f / area;
    if (isTextureActivated) {
      inter.normalMapIndex = normalMapIndex;
      float w = 1 - u - v;
      inter.textPos.x = uv0.x * w + uv1.x * u + uv2.x * v;
      inter.textPos.y = uv0.y * w + uv1.y * u + uv2.y * v;
      inter.diffuseIndex = textureIndex;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// BVH.hpp
// ---------------------------------------------------------------------------------------------
struct BVHNode {
  const Object* obj = nullptr;
  BoundBox bound;
  BVHNode* left = nullptr;
  BVHNode* right = nullptr;
};

struct Texture {  // Texture.hpp:9-39
  int width = 0, height = 0;
  std::vector<V3> rgb;
  V3 getRGBat(float u, float v) const {
    if (width == 0 && height == 0) return V3();
    if (u > 0) u = u - (int)u;
    else
      u = 1 - (fabsf(u) - (int)fabsf(u));
    if (v > 0) v = v - (int)v;
    else
      v = 1 - (fabsf(v) - (int)fabsf(v));
    int x = u * width;
    int y = v * height;
    int index = y * width + x;
    if (index < 0) index = 0;
    if (index >= (int)rgb.size()) index = (int)rgb.size() - 1;
    return rgb.at(index);
  }
};

struct Scene {
  std::vector<std::unique_ptr<Object>> objList;
  std::vector<std::unique_ptr<BVHNode>> pool;
  BVHNode* root = nullptr;
  std::vector<const Object*> lightlist;
  std::vector<Texture> maps[4];
  TutuCamera cam;
  V3 bkgcolor;
  float eta = 1.f;

  BVHNode* newNode() {
    pool.emplace_back(new BVHNode());
    return pool.back().get();
  }

  BVHNode* recursiveBuild(std::vector<const Object*> objs) {  // BVH.hpp:47-123
    BVHNode* res = newNode();
    if (objs.size() == 0) return res;
    else if (objs.size() == 1) {
      res->bound = objs.at(0)->bound;
      res->obj = objs[0];
      return res;
    } else if (objs.size() == 2) {
      res->left = recursiveBuild({objs[0]});
      res->right = recursiveBuild({objs[1]});
      res->bound = Union(res->left->bound, res->right->bound);
      return res;
    } else {
      BoundBox unionBound = Union(objs[0]->bound, objs[1]->bound);
      for (size_t i = 2; i < objs.size(); i++) unionBound = Union(unionBound, objs[i]->bound);
      int longest = unionBound.maxExtent();
      switch (longest) {
        case 0:
          std::sort(objs.begin(), objs.end(), [](const Object* o1, const Object* o2) -> bool {
            return o1->bound.Centroid().x < o2->bound.Centroid().x;
          });
          break;
        case 1:
          std::sort(objs.begin(), objs.end(), [](const Object* o1, const Object* o2) -> bool {
            return o1->bound.Centroid().y < o2->bound.Centroid().y;
          });
          break;
        case 2:
          std::sort(objs.begin(), objs.end(), [](const Object* o1, const Object* o2) -> bool {
            return o1->bound.Centroid().z < o2->bound.Centroid().z;
          });
          break;
      }
      auto begin = objs.begin();
      auto middle = begin + (objs.size() / 2);
      auto end = objs.end();
      std::vector<const Object*> leftObjects(begin, middle);
      std::vector<const Object*> rightObjects(middle, end);
      res->left = recursiveBuild(leftObjects);
      res->right = recursiveBuild(rightObjects);
      res->bound = Union(res->left->bound, res->right->bound);
    }
    return res;
  }
};

Intersection getIntersection(const BVHNode* node, const V3& rayOrig, const V3& rayDir) {  // BVH.hpp:145-167
  Intersection inter;
  if (!node) return inter;
  if (!node->bound.IntersectRay(rayOrig, rayDir)) return inter;
  if (!node->left && !node->right) {
    node->obj->intersect(rayOrig, rayDir, inter);
    return inter;
  }
  Intersection linter = getIntersection(node->left, rayOrig, rayDir);
  Intersection rinter = getIntersection(node->right, rayOrig, rayDir);
  if (linter.t <= rinter.t) return linter;
  return rinter;
}

bool hasIntersection(const BVHNode* node, const V3& rayOrig, const V3& rayDir, float dis) {  // BVH.hpp:170-194
  if (!node) return false;
  if (!node->bound.IntersectRay(rayOrig, rayDir)) return false;
  if (!node->left && !node->right) {
    Intersection inter;
    node->obj->intersect(rayOrig, rayDir, inter);
    if (inter.intersected && inter.t < dis && !FLOAT_EQUAL(inter.t, dis)) return true;
    return false;
  }
  if (hasIntersection(node->left, rayOrig, rayDir, dis)) return true;
  return hasIntersection(node->right, rayOrig, rayDir, dis);
}

// ---------------------------------------------------------------------------------------------
// IIntegrator.hpp helpers
// ---------------------------------------------------------------------------------------------
void changeNormalDir(Intersection& inter, const Scene& g) {  // :27-87
  const Texture& nMap = g.maps[TUTU_TEX_NORMAL].at(inter.normalMapIndex);
  V3 color = nMap.getRGBat(inter.textPos.x, inter.textPos.y);
  if (inter.obj->objectType == TUTU_PRIM_TRIANGLE) {
    const Object* t = inter.obj;
    V3 e1 = t->v1 - t->v0;
    V3 e2 = t->v2 - t->v0;
    V3 nDir = inter.Ns;
    nDir = normalized(nDir);
    float deltaU1 = t->uv1.x - t->uv0.x;
    float deltaV1 = t->uv1.y - t->uv0.y;
    float deltaU2 = t->uv2.x - t->uv0.x;
    float deltaV2 = t->uv2.y - t->uv0.y;
    float coef = 1 / (-deltaU1 * deltaV2 + deltaV1 * deltaU2);
    V3 T = coef * (-deltaV2 * e1 + deltaV1 * e2);
    V3 B = coef * (-deltaU2 * e1 + deltaU1 * e2);
    T = normalized(T);
    B = normalized(B);
    V3 res;
    res.x = T.x * color.x + B.x * color.y + nDir.x * color.z;
    res.y = T.y * color.x + B.y * color.y + nDir.y * color.z;
    res.z = T.z * color.x + B.z * color.y + nDir.z * color.z;
    inter.Ns = normalized(res);
  } else {
    V3 nDir = inter.Ng;
    V3 T = V3(-nDir.y / sqrtf(nDir.x * nDir.x + nDir.y * nDir.y),
              nDir.x / sqrtf(nDir.x * nDir.x + nDir.y * nDir.y), 0);
    V3 B = crossProduct(nDir, T);
    V3 res;
    res.x = T.x * color.x + B.x * color.y + nDir.x * color.z;
    res.y = T.y * color.x + B.y * color.y + nDir.y * color.z;
    res.z = T.z * color.x + B.z * color.y + nDir.z * color.z;
    inter.Ns = normalized(res);
  }
}

void textureModify(Intersection& inter, const Scene& g) {  // :89-127 (index checks are done at load)
  if (inter.diffuseIndex != -1)
    inter.mtlcolor.diffuse = g.maps[TUTU_TEX_DIFFUSE].at(inter.diffuseIndex).getRGBat(inter.textPos.x, inter.textPos.y);
  if (inter.normalMapIndex != -1) changeNormalDir(inter, g);
  if (inter.roughnessMapIndex != -1)
    inter.mtlcolor.roughness = g.maps[TUTU_TEX_ROUGHNESS].at(inter.roughnessMapIndex).getRGBat(inter.textPos.x, inter.textPos.y).x;
  if (inter.metallicMapIndex != -1)
    inter.mtlcolor.metallic = g.maps[TUTU_TEX_METALLIC].at(inter.metallicMapIndex).getRGBat(inter.textPos.x, inter.textPos.y).x;
}

bool isShadowRayBlocked(V3 orig, const V3& lightPos, const Scene& g) {  // :135-153
  V3 raydir = normalized(lightPos - orig);
  float distance = (lightPos - orig).norm();
  return hasIntersection(g.root, orig, raydir, distance);
}

float getLightPdf(const Intersection& inter, const Scene& g) {  // :155-168
  if (!inter.intersected) return 0;
  int size = (int)g.lightlist.size();
  if (size == 0) return 0;
  if (!inter.obj->mtlcolor.hasEmission()) return 0;
  float area = inter.obj->getArea();
  return 1 / (size * area);
}

void sampleLight(Intersection& inter, float& pdf, const Scene& g, const Rng& rng, int depth) {  // :173-192
  int size = (int)g.lightlist.size();
  if (size == 0) {
    inter.intersected = false;
    pdf = 0;
    return;
  }
  int index = (int)(rng.get(depth, 0) * (size - 1) + 0.4999f);
  if (size == 1) index = 0;
  const Object* lightObject = g.lightlist.at(index);
  lightObject->samplePoint(inter, pdf, rng.get(depth, 1), rng.get(depth, 2));
  pdf = (1.f / (size * lightObject->getArea()));
}

// ---------------------------------------------------------------------------------------------
// PathTracing.hpp
// ---------------------------------------------------------------------------------------------
struct Tracer {
  const Scene& g;
  Rng rng;
  uint64_t closest_calls = 0, any_calls = 0, shade_calls = 0;

  Intersection UpdateInter(const V3& o, const V3& d) {  // BVHStrategy.hpp:8-11
    ++closest_calls;
    return getIntersection(g.root, o, d);
  }

  V3 calcForRefractive(const V3& origin, const V3& dir, Intersection& inter, int depth) {  // :80-134
    if (depth > O_MAX_DEPTH) return 0;
    V3 Ng = inter.Ng;
    V3 Ns = inter.Ns;
    V3 wo = -dir;
    float eta_i = g.eta;
    float eta_t = inter.mtlcolor.eta;
    V3 wi;
    auto [sampleSuccess, TIR] = inter.mtlcolor.sampleDirection(
        wo, inter.Ns, wi, eta_i, [&](int k) { return rng.get(depth, 3 + k); });
    (void)sampleSuccess;
    wi = normalized(wi);
    float pdf = inter.mtlcolor.pdf(wi, wo, inter.Ns, eta_i, eta_t);
    if (TIR) {
      wi = normalized(getReflectionDir(wo, Ns));
      pdf = 1;
      if (inter.mtlcolor.mType == TUTU_MAT_MICROFACET_T) {
        V3 interNg = Ng;
        V3 interNs = Ns;
        if (wo.dot(Ng) < 0) {
          interNg = -interNg;
          std::swap(eta_i, eta_t);
          interNs = -interNs;
        }
        V3 h = normalized(wo + wi);
        float cosTheta = fabsf(interNs.dot(h));
        wi = normalized(getReflectionDir(wo, h));
        pdf = 1 * D_ndf(h, interNs, inter.mtlcolor.roughness) * cosTheta / (4.f * wo.dot(h));
      }
    }
    V3 f_r = inter.mtlcolor.BxDF(wi, wo, Ng, Ns, eta_i, false, TIR);
    V3 rayOrig = inter.pos;
    float cos = 0;
    if (wi.dot(Ns) > 0) {
      rayOrig = rayOrig + Ns * O_EPSILON;
      cos = fabsf(Ng.dot(wi));
    } else {
      rayOrig = rayOrig - Ns * O_EPSILON;
      cos = fabsf((-Ng).dot(wi));
    }
    V3 Li = traceRay(rayOrig, wi, depth + 1, V3(1));
    if (pdf < O_MIN_DIVISOR) return 0;
    return Li * cos * f_r / pdf;
  }

  V3 traceRay(const V3& origin, const V3& dir, int depth, V3 tp, Intersection* nxtInter = nullptr) {  // :136-279
    if (depth > O_MAX_DEPTH) return 0;
    V3 sampleValue = 0;
    Intersection inter;
    if (nxtInter) inter = *nxtInter;
    else
      inter = UpdateInter(origin, dir);
    if (!inter.intersected) return g.bkgcolor;
    ++shade_calls;
    if (inter.mtlcolor.mType == TUTU_MAT_PERFECT_REFRACTIVE || inter.mtlcolor.mType == TUTU_MAT_MICROFACET_T)
      return calcForRefractive(origin, dir, inter, depth);
    if (inter.obj->isTextureActivated) textureModify(inter, g);
    if (inter.mtlcolor.mType == TUTU_MAT_UNLIT) return inter.mtlcolor.diffuse;
    if (inter.mtlcolor.hasEmission() && depth > 0) return 0;
    if (inter.mtlcolor.hasEmission()) return inter.mtlcolor.emission;
    V3 wo = -dir;

    float light_pdf;
    float mis_weight_l = 0.f;
    float mat_pdf;
    float mis_weight_m = 0.f;
    Intersection light_inter;
    sampleLight(light_inter, light_pdf, g, rng, depth);
    bool rayInside = inter.Ns.dot(wo) < 0;
    V3 shadowRayOrig = inter.pos;
    V3 lightPos = light_inter.pos;
    offsetRayOrig(shadowRayOrig, inter.Ns, rayInside);
    offsetRayOrig(lightPos, light_inter.Ns, false);
    bool blocked = true;
    if (light_inter.intersected) {
      ++any_calls;
      blocked = isShadowRayBlocked(shadowRayOrig, lightPos, g);
    }
    if (!light_inter.intersected || blocked) {
    } else {
      V3 wi = light_inter.pos - inter.pos;
      float r2 = wi.norm2();
      wi = normalized(wi);
      if (wi.dot(light_inter.Ns) > 0) {
      } else {
        mat_pdf = inter.mtlcolor.pdf(wi, wo, inter.Ns, g.eta, inter.mtlcolor.eta);
        V3 light_N = normalized(light_inter.Ns);
        float cos_theta_prime = light_N.dot(-wi);
        if (cos_theta_prime <= 0) goto jmp;
        float dot = inter.Ng.dot(wi);
        float cos_theta = fabsf(dot);
        float pdfl = light_pdf;
        light_pdf = light_pdf * r2 / cos_theta_prime;
        mis_weight_l = getMisWeight(light_pdf, mat_pdf);
        V3 f_r = inter.mtlcolor.BxDF(wi, wo, inter.Ng, inter.Ns, g.eta);
        V3 L_i = light_inter.mtlcolor.emission;
        if (r2 * pdfl < O_MIN_DIVISOR) return sampleValue;
        sampleValue = sampleValue + (mis_weight_l * L_i * f_r * cos_theta * cos_theta_prime / (r2 * pdfl));
      }
    }
  jmp:
    V3 wi;
    auto [sampleSucess, specialEvent] = inter.mtlcolor.sampleDirection(
        wo, inter.Ns, wi, g.eta, [&](int k) { return rng.get(depth, 3 + k); });
    (void)specialEvent;
    if (!sampleSucess) return sampleValue;
    mat_pdf = inter.mtlcolor.pdf(wi, wo, inter.Ns, g.eta, inter.mtlcolor.eta);
    Intersection x_inter;
    V3 rayOrig = inter.pos;
    offsetRayOrig(rayOrig, inter.Ns, wi.dot(inter.Ns) < 0);
    x_inter = UpdateInter(rayOrig, wi);
    if (!x_inter.intersected) {
    } else {
      float dot = fabsf(inter.Ng.dot(wi));
      float cos_theta = dot;
      light_pdf = getLightPdf(x_inter, g);
      if (light_pdf) {
        V3 light_N = normalized(x_inter.Ns);
        float cos_theta_prime = light_N.dot(-wi);
        if (cos_theta_prime <= 0) goto jmp2;
        float r2 = (x_inter.pos - inter.pos).norm2();
        float l_pdf_transformed = light_pdf * r2 / cos_theta_prime;
        mis_weight_m = getMisWeight(mat_pdf, l_pdf_transformed);
        if (inter.mtlcolor.mType == TUTU_MAT_PERFECT_REFLECTIVE && mat_pdf == 1.f) mis_weight_m = 1.f;
        V3 f_r = inter.mtlcolor.BxDF(wi, wo, inter.Ng, inter.Ns, g.eta);
        V3 L_i = x_inter.mtlcolor.emission;
        if (mat_pdf < O_MIN_DIVISOR) return sampleValue;
        sampleValue = sampleValue + (mis_weight_m * L_i * f_r * cos_theta / mat_pdf);
        return sampleValue;
      } else {
      jmp2:
        tp = depth > O_MIN_DEPTH ? tp : V3(1);
        float rr_prob = std::max(tp.x, std::max(tp.y, tp.z));
        if (rng.get(depth, 5) > rr_prob) return sampleValue;
        V3 f_r = inter.mtlcolor.BxDF(wi, wo, inter.Ng, inter.Ns, g.eta);
        V3 coe = f_r * cos_theta / (mat_pdf * rr_prob);
        if (mat_pdf * rr_prob < O_MIN_DIVISOR) return sampleValue;
        tp = tp * coe;
        V3 Li = traceRay(rayOrig, wi, depth + 1, tp, &x_inter);
        sampleValue = sampleValue + (Li * coe);
      }
    }
    return sampleValue;
  }
};

struct RayGenK {
  V3 eye, ul, delta_h, delta_v, c_off_h, c_off_v;
};

RayGenK raygen_constants(const TutuCamera& c) {  // Camera.hpp:12-48 + PathTracing.hpp:357-391
  V3 fwd = normalized(V3(c.viewdir[0], c.viewdir[1], c.viewdir[2]));
  V3 upIn(c.updir[0], c.updir[1], c.updir[2]);
  V3 right = normalized(crossProduct(fwd, upIn));
  V3 up = normalized(crossProduct(right, fwd));
  float tanHalfHfov = tanf((c.hfov_deg * 0.5f) * O_PI / 180.f);
  float imagePlaneDist = c.width / (2.f * tanHalfHfov);
  V3 u = normalized(crossProduct(fwd, up));
  V3 v = normalized(crossProduct(u, fwd));
  float d = imagePlaneDist;
  if (c.parallel_projection) d = 4.f;
  float width_half = fabsf(tanf((c.hfov_deg / 2.f) * O_PI / 180.f) * d);
  float aspect_ratio = c.width / (float)c.height;
  float height_half = width_half / aspect_ratio;
  V3 n = normalized(V3(c.viewdir[0], c.viewdir[1], c.viewdir[2]));
  RayGenK k;
  k.eye = V3(c.eye[0], c.eye[1], c.eye[2]);
  k.ul = k.eye + d * n - width_half * u + height_half * v;
  V3 ur = k.eye + d * n + width_half * u + height_half * v;
  V3 ll = k.eye + d * n - width_half * u - height_half * v;
  k.delta_h = V3(0, 0, 0);
  if (c.width != 1) k.delta_h = (ur - k.ul) / (float)(c.width - 1);
  k.delta_v = V3(0, 0, 0);
  if (c.height != 1) k.delta_v = (ll - k.ul) / (float)(c.height - 1);
  k.c_off_h = (ur - k.ul) / (float)(c.width * 2);
  k.c_off_v = (ll - k.ul) / (float)(c.height * 2);
  return k;
}

#include "tutu_oracle_bdpt.hpp"

}  // namespace

// =============================================================================================
// C API (ctypes)
// =============================================================================================
struct OracleScene {
  Scene s;
};

extern "C" {

OracleScene* oracle_scene_create(const TutuSceneDesc* d) {
  if (!d || d->struct_size != sizeof(TutuSceneDesc)) return nullptr;
  std::unique_ptr<OracleScene> os(new OracleScene());
  Scene& s = os->s;
  s.cam = d->camera;
  s.bkgcolor = V3(d->bkgcolor[0], d->bkgcolor[1], d->bkgcolor[2]);
  s.eta = d->eta;
  for (int c = 0; c < 4; ++c)
    for (uint32_t i = 0; i < d->n_tex[c]; ++i) {
      Texture t;
      t.width = d->tex[c][i].width;
      t.height = d->tex[c][i].height;
      size_t n = (size_t)t.width * t.height;
      t.rgb.resize(n);
      for (size_t k = 0; k < n; ++k)
        t.rgb[k] = V3(d->tex[c][i].rgb[3 * k], d->tex[c][i].rgb[3 * k + 1], d->tex[c][i].rgb[3 * k + 2]);
      s.maps[c].push_back(std::move(t));
    }
  for (uint32_t i = 0; i < d->n_prims; ++i) {
    const TutuPrim& p = d->prims[i];
    std::unique_ptr<Object> o(new Object());
    o->objectType = p.type;
    o->index = (int)i;
    const TutuMaterial& m = d->materials[p.material];
    o->mtlcolor.diffuse = V3(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
    o->mtlcolor.specular = V3(m.specular[0], m.specular[1], m.specular[2]);
    o->mtlcolor.emission = V3(m.emission[0], m.emission[1], m.emission[2]);
    o->mtlcolor.mType = m.type;
    o->mtlcolor.alpha = m.alpha;
    o->mtlcolor.eta = m.eta;
    o->mtlcolor.roughness = m.roughness;
    o->mtlcolor.metallic = m.metallic;
    o->isTextureActivated = p.tex_active != 0;
    o->textureIndex = p.tex_diffuse;
    o->normalMapIndex = p.tex_normal;
    o->roughnessMapIndex = p.tex_roughness;
    o->metallicMapIndex = p.tex_metallic;
    if (p.type == TUTU_PRIM_SPHERE) {
      o->centerPos = V3(p.v[0], p.v[1], p.v[2]);
      o->radius = p.v[3];
    } else {
      o->v0 = V3(p.v[0], p.v[1], p.v[2]);
      o->v1 = V3(p.v[3], p.v[4], p.v[5]);
      o->v2 = V3(p.v[6], p.v[7], p.v[8]);
      o->n0 = V3(p.n[0], p.n[1], p.n[2]);
      o->n1 = V3(p.n[3], p.n[4], p.n[5]);
      o->n2 = V3(p.n[6], p.n[7], p.n[8]);
      o->uv0.x = p.uv[0], o->uv0.y = p.uv[1];
      o->uv1.x = p.uv[2], o->uv1.y = p.uv[3];
      o->uv2.x = p.uv[4], o->uv2.y = p.uv[5];
    }
    o->initializeBound();
    s.objList.push_back(std::move(o));
  }
  // Scene::initializeBVH (Scene.hpp:28-35) — always the oracle's own build, never desc->bvh_nodes
  std::vector<const Object*> objl;
  for (auto& o : s.objList) objl.push_back(o.get());
  if (!objl.empty()) s.root = s.recursiveBuild(objl);
  // PPMGenerator::initializeLights (PPMGenerator.hpp:317-324)
  for (auto& o : s.objList)
    if (o->mtlcolor.hasEmission()) s.lightlist.push_back(o.get());
  return os.release();
}

void oracle_scene_destroy(OracleScene* s) { delete s; }

uint32_t oracle_bvh_node_count(const OracleScene* os) { return (uint32_t)os->s.pool.size(); }

// pre-order export, same format as TutuBvhNode
uint32_t oracle_bvh_export(const OracleScene* os, TutuBvhNode* out) {
  if (!os->s.root) return 0;
  struct Item {
    const BVHNode* n;
    int32_t parent;
    bool is_right;
  };
  std::vector<Item> st;
  st.push_back({os->s.root, -1, false});
  uint32_t count = 0;
  while (!st.empty()) {
    Item it = st.back();
    st.pop_back();
    int32_t me = (int32_t)count++;
    bool leaf = !it.n->left && !it.n->right;
    out[me] = {-1, -1, leaf ? it.n->obj->index : -1};
    if (it.parent >= 0) (it.is_right ? out[it.parent].right : out[it.parent].left) = me;
    if (!leaf) {
      st.push_back({it.n->right, me, true});
      st.push_back({it.n->left, me, false});
    }
  }
  return count;
}

void oracle_trace_closest(const OracleScene* os, const float* rays, uint64_t n, TutuHit* out, int threads) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([=] {
      for (uint64_t i = n * w / threads; i < n * (w + 1) / threads; ++i) {
        const float* r = rays + i * TUTU_RAY_FLOATS;
        Intersection it = getIntersection(os->s.root, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]));
        out[i].prim = it.intersected ? it.obj->index : -1;
        out[i].t = it.t;
        out[i].u = it.intersected ? it.u : 0.f;
        out[i].v = it.intersected ? it.v : 0.f;
      }
    });
  for (auto& t : pool) t.join();
}

void oracle_trace_any(const OracleScene* os, const float* rays, uint64_t n, uint8_t* out, int threads) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([=] {
      for (uint64_t i = n * w / threads; i < n * (w + 1) / threads; ++i) {
        const float* r = rays + i * TUTU_RAY_FLOATS;
        out[i] = hasIntersection(os->s.root, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[7]) ? 1 : 0;
      }
    });
  for (auto& t : pool) t.join();
}

// PathTracing::integrate + sub_render_pt (PathTracing.hpp:352-475,485-516): samples
// [sample_begin, sample_begin+sample_count) of every pixel, NaN samples dropped, result
// = sum * (1/total_spp) written to rgb_out[(y*W+x)*3].  counters_out (optional, 3 x uint64):
// closest-hit calls, any-hit calls, shading-vertex evaluations.
void oracle_render_path(const OracleScene* os, uint32_t sample_begin, uint32_t sample_count,
                        uint32_t total_spp, uint64_t seed, float* rgb_out, uint64_t* counters_out,
                        int threads) {
  const Scene& s = os->s;
  const int W = s.cam.width, H = s.cam.height;
  RayGenK k = raygen_constants(s.cam);
  const float SPP_inv = 1.f / total_spp;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::atomic<int> next{0};
  std::atomic<uint64_t> c0{0}, c1{0}, c2{0};
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([&] {
      Tracer tr{s, Rng{seed, 0, 0, true}};
      for (;;) {
        int y = next.fetch_add(1);
        if (y >= H) break;
        for (int x = 0; x < W; ++x) {
          V3 pixelPos = k.ul + (float)x * k.delta_h + (float)y * k.delta_v + k.c_off_v + k.c_off_v;  // :503 (sic)
          V3 rayDir = normalized((pixelPos - k.eye));
          V3 estimate;
          for (uint32_t i = 0; i < sample_count; ++i) {
            tr.rng.pixel = (uint32_t)(y * W + x);
            tr.rng.sample = sample_begin + i;
            V3 res = tr.traceRay(k.eye, rayDir, 0, V3(1), nullptr);
            if (!std::isnan(res.x) && !std::isnan(res.y) && !std::isnan(res.z)) estimate = estimate + res;
          }
          V3 color = estimate * SPP_inv;
          float* o = rgb_out + ((size_t)y * W + x) * 3;
          o[0] = color.x, o[1] = color.y, o[2] = color.z;
        }
      }
      c0 += tr.closest_calls;
      c1 += tr.any_calls;
      c2 += tr.shade_calls;
    });
  for (auto& t : pool) t.join();
  if (counters_out) {
    counters_out[0] = c0;
    counters_out[1] = c1;
    counters_out[2] = c2;
  }
}

// BDPT::integrate (BDPT.hpp:395-674 -> sub_render_bdpt :679-900).  The frame buffer starts at
// bkgcolor (Camera.hpp:28) and contributions are ADDED (:891); t == 1 strategies splat into other
// pixels (:820-824), collected per thread here and summed in thread order at the end.
// counters_out: closest-hit calls, any-hit calls, general connections attempted.
void oracle_render_bdpt(const OracleScene* os, uint32_t sample_begin, uint32_t sample_count, uint32_t total_spp,
                        uint64_t seed, float* rgb_out, uint64_t* counters_out, int threads) {
  const Scene& s = os->s;
  const int W = s.cam.width, H = s.cam.height;
  RayGenK k = raygen_constants(s.cam);
  const Cam cam = make_cam(s.cam);
  const float SPP_inv = 1.f / total_spp;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads <= 0) threads = 1;
  std::atomic<int> next{0};
  std::atomic<uint64_t> c0{0}, c1{0}, c2{0};
  std::vector<std::vector<float>> splats(threads, std::vector<float>((size_t)W * H * 3, 0.f));
  for (size_t i = 0; i < (size_t)W * H; ++i) {
    rgb_out[3 * i] = s.bkgcolor.x, rgb_out[3 * i + 1] = s.bkgcolor.y, rgb_out[3 * i + 2] = s.bkgcolor.z;
  }
  std::vector<std::thread> pool;
  for (int w = 0; w < threads; ++w)
    pool.emplace_back([&, w] {
      Bdpt b{s, cam, Rng{seed, 0, 0}};
      std::vector<float>& mine = splats[w];
      for (;;) {
        int y = next.fetch_add(1);
        if (y >= H) break;
        for (int x = 0; x < W; ++x) {
          V3 pixelPos = k.ul + (float)x * k.delta_h + (float)y * k.delta_v + k.c_off_h + k.c_off_v;  // :700
          V3 rayDir = normalized(pixelPos - k.eye);
          V3 estimate;
          for (uint32_t i = 0; i < sample_count; ++i) {
            b.rng.pixel = (uint32_t)(y * W + x);
            b.rng.sample = sample_begin + i;
            if (!b.sample(k.eye, pixelPos, rayDir, SPP_inv, estimate, [&](int index, const V3& v) {
                  mine[3 * (size_t)index] += v.x, mine[3 * (size_t)index + 1] += v.y, mine[3 * (size_t)index + 2] += v.z;
                }))
              break;
          }
          V3 add = estimate * SPP_inv;  // :891
          float* o = rgb_out + ((size_t)y * W + x) * 3;
          o[0] += add.x, o[1] += add.y, o[2] += add.z;
        }
      }
      c0 += b.closest_calls;
      c1 += b.any_calls;
      c2 += b.connections;
    });
  for (auto& t : pool) t.join();
  for (int w = 0; w < threads; ++w)
    for (size_t i = 0; i < (size_t)W * H * 3; ++i) rgb_out[i] += splats[w][i];
  if (counters_out) {
    counters_out[0] = c0;
    counters_out[1] = c1;
    counters_out[2] = c2;
  }
}

// PPMGenerator::writePixel (PPMGenerator.hpp:812-845) with GAMMA_COORECTION (global.hpp:29-30):
// color = 255 * pow(clamp(0, 1, color), 0.78f), written as (int)color.  `pow` on two floats is the
// float overload (powf); clamp is std::max(lo, std::min(hi, v)) (global.hpp:52-55), so NaN -> 1.
void oracle_write_pixel(const float* rgb, uint64_t n_values, float gamma, uint8_t* out) {
  for (uint64_t i = 0; i < n_values; ++i) {
    float c = clampf(0.f, 1.f, rgb[i]);
    float v = gamma > 0.f ? 255 * powf(c, gamma) : 255 * c;
    out[i] = (uint8_t)(int)v;
  }
}

// ---------------------------------------------------------------------------------------------
// Postprocessor (Postprocessor.hpp:29-197) restated.  mode: 1 getEmmisiveTexture (:131-156), 2
// getGaussianBlurTexture (:64-128), 3 the bloom chain of performPostProcess (:37-50), 4 getHDRtexture
// (:182-207), 5 bloom then tone map (performPostProcess under HDR_BLOOM).  Constants are the reference's
// #defines (:10-14); every texel read goes through Texture::getRGBat (Texture.hpp:18-39) as there.
// ---------------------------------------------------------------------------------------------
namespace pp {
struct Tex {
  int width = 0, height = 0;
  std::vector<V3> rgb;
  V3 getRGBat(float u, float v) const {  // Texture.hpp:18-39
    if (width == 0 && height == 0) return V3();
    if (u > 0)
      u = u - (int)u;
    else
      u = 1 - (std::fabs(u) - (int)std::fabs(u));
    if (v > 0)
      v = v - (int)v;
    else
      v = 1 - (std::fabs(v) - (int)std::fabs(v));
    int x = u * width;
    int y = v * height;
    long long index = (long long)y * width + x;
    if (index < 0) index = 0;
    if (index >= (long long)rgb.size()) index = (long long)rgb.size() - 1;
    return rgb[(size_t)index];
  }
};
Tex like(const Tex& s) {
  Tex r;
  r.width = s.width, r.height = s.height;
  r.rgb.assign((size_t)s.width * s.height, V3());
  return r;
}
float rescale(float input, float originMax, float originMin, float targetMax, float targetMin) {  // global.hpp:66-68
  return targetMin + ((targetMax - targetMin) * (input - originMin) / (originMax - originMin));
}
Tex emissive(const Tex& src, float threshold, float strength) {  // :131-156
  Tex res = like(src);
  int h = res.height, w = res.width;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      V3& color = res.rgb[(size_t)y * w + x];
      float U = (float)x / w, V = (float)y / h;
      V3 col = src.getRGBat(clampf(0, 0.999f, U), clampf(0, 0.999f, V));
      if (sqrtf(col.x * col.x + col.y * col.y + col.z * col.z) > threshold) {
        float mx = col.x > col.y ? col.x : col.y;
        mx = mx > col.z ? mx : col.z;
        color.x = rescale(col.x, mx, 0.f, strength, 0.f);
        color.y = rescale(col.y, mx, 0.f, strength, 0.f);
        color.z = rescale(col.z, mx, 0.f, strength, 0.f);
      }
    }
  return res;
}
Tex blur(const Tex& img, int kernelSize, float stddev) {  // :64-128
  Tex src = img;
  Tex res = like(src);
  int h = res.height, w = res.width;
  const float E = 2.7182818f;
  const float PI_F = 3.1415926535897f;  // global.hpp:15
  auto gaussian = [&](int inputX, float standardDev) -> float {
    return (1 / sqrtf(2 * PI_F * standardDev)) * powf(E, -(inputX * inputX) / (2 * standardDev * standardDev));
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      V3& color = res.rgb[(size_t)y * w + x];
      int startY = -kernelSize * 0.5;
      V3 col;
      float kernelSum = 0;
      for (int i = 0; i < kernelSize; i++) {
        float U = (float)x / w;
        float V = (float)(y + i + startY) / h;
        float gauss = gaussian(startY + i, stddev);
        col = col + src.getRGBat(clampf(0, 0.999f, U), clampf(0, 0.999f, V)) * gauss;
        kernelSum += gauss;
      }
      color = V3(col.x / kernelSum, col.y / kernelSum, col.z / kernelSum);
    }
  src = res;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      V3& color = res.rgb[(size_t)y * w + x];
      int startX = -kernelSize * 0.5;
      V3 col;
      float kernelSum = 0;
      for (int i = 0; i < kernelSize; i++) {
        float U = (float)(x + i + startX) / w;
        float V = (float)y / h;
        float gauss = gaussian(startX + i, stddev);
        col = col + src.getRGBat(clampf(0, 0.999f, U), clampf(0, 0.999f, V)) * gauss;
        kernelSum += gauss;
      }
      color = V3(col.x / kernelSum, col.y / kernelSum, col.z / kernelSum);
    }
  return res;
}
Tex add(const Tex& a, const Tex& b) {  // :158-175
  Tex res = a;
  for (size_t i = 0; i < res.rgb.size(); ++i) res.rgb[i] = V3(res.rgb[i].x + b.rgb[i].x, res.rgb[i].y + b.rgb[i].y, res.rgb[i].z + b.rgb[i].z);
  return res;
}
Tex hdr(const Tex& src, float exposure) {  // :182-207
  Tex dst = like(src);
  for (int y = 0; y < dst.height; y++)
    for (int x = 0; x < dst.width; x++) {
      float U = (float)x / dst.width, V = (float)y / dst.height;
      V3 c = src.getRGBat(clampf(0, 0.999f, U), clampf(0, 0.999f, V));
      dst.rgb[(size_t)y * src.width + x] = V3(1 - expf(-c.x * exposure), 1 - expf(-c.y * exposure), 1 - expf(-c.z * exposure));
    }
  return dst;
}
}  // namespace pp

int oracle_postprocess(const float* rgb, int width, int height, int mode, const TutuPostParams* params, float* out) {
  TutuPostParams p{3.f, 2.f, 1, 10, 30.f, 1.5f};  // Postprocessor.hpp:10-14, :141
  if (params) p = *params;
  pp::Tex src;
  src.width = width, src.height = height;
  src.rgb.resize((size_t)width * height);
  for (size_t i = 0; i < src.rgb.size(); ++i) src.rgb[i] = V3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
  pp::Tex res;
  if (mode == TUTU_POST_EXTRACT) {
    res = pp::emissive(src, p.emissive_norm, p.strength);
  } else if (mode == TUTU_POST_BLUR) {
    res = pp::blur(src, p.kernel_size, p.stddev);
  } else if (mode == TUTU_POST_HDR) {
    res = pp::hdr(src, p.exposure);
  } else if (mode == TUTU_POST_BLOOM || mode == TUTU_POST_HDR_BLOOM) {
    pp::Tex e = pp::emissive(src, p.emissive_norm, p.strength);
    pp::Tex b = pp::blur(e, p.kernel_size, p.stddev);
    for (int i = 0; i < p.gaussian_loops; ++i) b = pp::blur(b, p.kernel_size, p.stddev);
    res = pp::add(src, b);
    if (mode == TUTU_POST_HDR_BLOOM) res = pp::hdr(res, p.exposure);
  } else {
    return -1;
  }
  for (size_t i = 0; i < res.rgb.size(); ++i) out[3 * i] = res.rgb[i].x, out[3 * i + 1] = res.rgb[i].y, out[3 * i + 2] = res.rgb[i].z;
  return 0;
}

// primary rays exactly as sub_render_pt generates them (for ray-generation parity tests)
void oracle_primary_rays(const OracleScene* os, float* rays_out) {
  const Scene& s = os->s;
  RayGenK k = raygen_constants(s.cam);
  for (int y = 0; y < s.cam.height; ++y)
    for (int x = 0; x < s.cam.width; ++x) {
      V3 pixelPos = k.ul + (float)x * k.delta_h + (float)y * k.delta_v + k.c_off_v + k.c_off_v;
      V3 d = normalized((pixelPos - k.eye));
      float* o = rays_out + ((size_t)y * s.cam.width + x) * TUTU_RAY_FLOATS;
      o[0] = k.eye.x, o[1] = k.eye.y, o[2] = k.eye.z, o[3] = 0;
      o[4] = d.x, o[5] = d.y, o[6] = d.z, o[7] = FLT_MAX;
    }
}

}  // extern "C"
