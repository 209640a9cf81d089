// trace.cuh — BVH closest-hit / any-hit traversal and primitive intersection for sm_100a.
//
// Arithmetic contract: everything that decides WHICH primitive is hit is written with explicit
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn), which ptxas
// never contracts into FMAs, in the operand order of the reference's scalar C++:
//   BoundBox::IntersectRay  reference include/BoundBox.hpp:55-92
//   Triangle::intersect     reference include/Triangle.hpp:23-74
//   Sphere::intersect       reference include/Sphere.hpp:26-126 (+ solveQuadratic global.hpp:147-167)
//   getIntersection         reference include/BVH.hpp:145-167   (tie: leftmost DFS leaf wins)
//   hasIntersection         reference include/BVH.hpp:170-194
// B200 has no RT cores: traversal is software, one ray per thread, a short per-thread stack,
// one 64-byte inner node (both child boxes) fetched as 4 x LDG.128 through L1/L2.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace tutu {

constexpr int kStackSize = 32;          // >= tree depth; checked at upload
constexpr uint32_t kSphereBit = 1u << 30;
constexpr uint32_t kSlotMask = kSphereBit - 1u;

struct DevScene {
  const float4* __restrict__ inner;     // 4 x float4 per inner node, the reference's topology
  const float4* __restrict__ inner_fast;  // SAH topology over the same leaves: regular rays (host_scene.cpp)
  const float4* __restrict__ wide;      // compressed 8-wide tree for regular rays: 6 x 16 B per node (wide.cuh); nullptr = none
  const float4* __restrict__ wleaf;     // its leaf records: 4 x float4 (LeafGeom + leaf code), in wide-tree order
  const float4* __restrict__ wbox;      // exact leaf boxes, 2 x float4, same order
  const float4* __restrict__ geom;      // 3 x float4 per leaf slot
  const float4* __restrict__ shade;     // 4 x float4 per leaf slot
  const int4* __restrict__ leaftex;     // per leaf slot, or nullptr when no prim is textured
  const int* __restrict__ slot_to_prim;
  const float4* __restrict__ materials; // 4 x float4 per material
  const float4* __restrict__ lights;    // 8 x float4 per light
  const int4* __restrict__ tex_headers[4];  // {width, height, offset, n_texels}
  const float4* __restrict__ texels;
  float root_lo[3];
  float root_hi[3];
  int root_ref;
  int root_ref_fast;
  int empty;
  int n_lights;
  float bkg[3];
  float eta;
  // Pruning slack (see traverse()): a subtree is culled only when its box is entered later than
  // best_t * (1 + prune_rel) + prune_abs.
  float prune_rel;
  float prune_abs;
  int refill_min;  // persistent tracer: refill once this many lanes of a warp are idle
  int leaf_batch;  // traverse_batched: waiting lanes that trigger the primitive tests
  int queue_lanes; // BDPT queue tracers: 1 = persistent lanes with phase-separated steps (trace_queue_lanes), 0 = one packet at a time
  int lanes_leaf_batch;  // trace_queue_lanes: waiting lanes that trigger the primitive tests
  unsigned sphere_mask;  // small scenes: bit s set = leaf slot s is a sphere
};

struct Ray {
  float ox, oy, oz;
  float dx, dy, dz;
};

struct Hit {
  float t, u, v;
  int slot;  // leaf slot (DFS order) | kSphereBit for spheres, -1 = miss
};

// ---- exact float helpers (Vector.hpp:186,213-225) ------------------------------------------
__device__ __forceinline__ float dot_rn(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
__device__ __forceinline__ float cross_c(float a, float b, float c, float d) {  // a*b - c*d
  return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
// normalized(): mag = sqrtf(x*x+y*y+z*z); if (mag>0) v * (1/mag) else v
__device__ __forceinline__ void normalize_rn(float& x, float& y, float& z) {
  float mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  if (mag > 0.f) {
    float inv = __frcp_rn(mag);
    x = __fmul_rn(x, inv);
    y = __fmul_rn(y, inv);
    z = __fmul_rn(z, inv);
  }
}

// __frcp_rn(x) is the correctly rounded 1/x, i.e. the same float as the reference's `1 / x` (IEEE division), in about half
// the instructions of __fdiv_rn(1.f, x).
struct RayPre {  // per-ray constants of the slab test
  float ox, oy, oz;
  float ix, iy, iz;     // 1/d, IEEE (inf for 0)
  bool nx, ny, nz;      // d < 0 (false for -0, like the reference's std::swap condition)
};

__device__ __forceinline__ RayPre make_pre(const Ray& r) {
  RayPre p;
  p.ox = r.ox, p.oy = r.oy, p.oz = r.oz;
  p.ix = __frcp_rn(r.dx);
  p.iy = __frcp_rn(r.dy);
  p.iz = __frcp_rn(r.dz);
  p.nx = r.dx < 0.f;
  p.ny = r.dy < 0.f;
  p.nz = r.dz < 0.f;
  return p;
}

// BoundBox::IntersectRay.  The reference computes tmin/tmax per axis and swaps them when d<0;
// selecting the plane before the (identical) subtract-multiply gives the same two values.
// The ternaries are kept literally: with NaN operands (0*inf) they are NOT fmaxf/fminf.
__device__ __forceinline__ bool box_test(const RayPre& p, float lox, float loy, float loz, float hix,
                                         float hiy, float hiz, float& t_enter) {
  float tmin_x = __fmul_rn(__fsub_rn(p.nx ? hix : lox, p.ox), p.ix);
  float tmax_x = __fmul_rn(__fsub_rn(p.nx ? lox : hix, p.ox), p.ix);
  float tmin_y = __fmul_rn(__fsub_rn(p.ny ? hiy : loy, p.oy), p.iy);
  float tmax_y = __fmul_rn(__fsub_rn(p.ny ? loy : hiy, p.oy), p.iy);
  float tmin_z = __fmul_rn(__fsub_rn(p.nz ? hiz : loz, p.oz), p.iz);
  float tmax_z = __fmul_rn(__fsub_rn(p.nz ? loz : hiz, p.oz), p.iz);
  float buffer = tmin_y > tmin_z ? tmin_y : tmin_z;
  t_enter = tmin_x > buffer ? tmin_x : buffer;
  buffer = tmax_y < tmax_z ? tmax_y : tmax_z;
  float t_exit = tmax_x < buffer ? tmax_x : buffer;
  return t_enter <= t_exit && t_exit >= 0.f;
}

// Triangle::intersect with E1, E2 and the unit normal precomputed on the host (same float ops).
__device__ __forceinline__ bool tri_test(const float4* __restrict__ g, const Ray& r, float& t, float& u,
                                         float& v) {
  const float4 a = __ldg(g + 0);  // v0.xyz, E1.x
  const float4 b = __ldg(g + 1);  // E1.y, E1.z, E2.x, E2.y
  const float4 c = __ldg(g + 2);  // E2.z, n.xyz
  const float e1x = a.w, e1y = b.x, e1z = b.y;
  const float e2x = b.z, e2y = b.w, e2z = c.x;
  // FLOAT_EQUAL(dir.dot(normal), 0.f): fabs(x - 0) < 1e-4
  const float dn = dot_rn(r.dx, r.dy, r.dz, c.y, c.z, c.w);
  if (fabsf(dn) < 0.0001f) return false;
  const float sx = __fsub_rn(r.ox, a.x), sy = __fsub_rn(r.oy, a.y), sz = __fsub_rn(r.oz, a.z);
  // S1 = dir x E2, S2 = S x E1
  const float s1x = cross_c(r.dy, e2z, r.dz, e2y);
  const float s1y = cross_c(r.dz, e2x, r.dx, e2z);
  const float s1z = cross_c(r.dx, e2y, r.dy, e2x);
  const float det = dot_rn(s1x, s1y, s1z, e1x, e1y, e1z);
  if (det == 0.f) return false;
  const float s2x = cross_c(sy, e1z, sz, e1y);
  const float s2y = cross_c(sz, e1x, sx, e1z);
  const float s2z = cross_c(sx, e1y, sy, e1x);
  const float left = __frcp_rn(det);
  t = __fmul_rn(left, dot_rn(s2x, s2y, s2z, e2x, e2y, e2z));
  u = __fmul_rn(left, dot_rn(s1x, s1y, s1z, sx, sy, sz));
  v = __fmul_rn(left, dot_rn(s2x, s2y, s2z, r.dx, r.dy, r.dz));
  return t > 0.f && __fsub_rn(__fsub_rn(1.f, u), v) > 0.f && u > 0.f && v > 0.f;
}

// Sphere::intersect.  C is summed in double (std::pow(float,int) promotes) and rounded once.
struct SphereHit {
  bool hit;
  float t;
};
// by-value arguments and result: a __noinline__ callee taking references would force the caller's
// ray and walk state into local memory
static __device__ __noinline__ SphereHit sphere_test(const float4* __restrict__ g, const Ray r) {
  float t = 0.f;
  const float4 a = __ldg(g + 0);  // centre.xyz, radius
  const float ocx = __fsub_rn(r.ox, a.x), ocy = __fsub_rn(r.oy, a.y), ocz = __fsub_rn(r.oz, a.z);
  const float B = __fmul_rn(2.f, dot_rn(r.dx, r.dy, r.dz, ocx, ocy, ocz));
  const double Cd = __dsub_rn(
      __dadd_rn(__dadd_rn(__dmul_rn((double)ocx, (double)ocx), __dmul_rn((double)ocy, (double)ocy)),
                __dmul_rn((double)ocz, (double)ocz)),
      (double)__fmul_rn(a.w, a.w));
  const float C = __double2float_rn(Cd);
  // solveQuadratic with A = 1: B*B - 4*A*C
  const float disc = __fsub_rn(__fmul_rn(B, B), __fmul_rn(4.f, C));
  float t1, t2;
  if (disc < 0.f) {
    t1 = FLT_MAX;
    t2 = FLT_MAX;
  } else if (disc == 0.f) {
    t1 = __fdiv_rn(__fadd_rn(-B, __fsqrt_rn(disc)), 2.f);
    t2 = t1;
  } else {
    t1 = __fdiv_rn(__fadd_rn(-B, __fsqrt_rn(disc)), 2.f);
    t2 = __fdiv_rn(__fsub_rn(-B, __fsqrt_rn(disc)), 2.f);
  }
  if (t1 > t2) {
    float s = t1;
    t1 = t2;
    t2 = s;
  }
  const bool eq1 = fabsf(__fsub_rn(t1, FLT_MAX)) < 0.0001f;
  const bool eq2 = fabsf(__fsub_rn(t2, FLT_MAX)) < 0.0001f;
  if (eq1 && eq2) return SphereHit{false, 0.f};
  if (fabsf(__fsub_rn(t1, t2)) < 0.0001f) {  // "one solution"
    if (t1 < 0.f) return SphereHit{false, 0.f};
    return SphereHit{true, t1};
  }
  if (t1 > 0.f && t2 > 0.f) t = t1;
  else if (t1 > 0.f && t2 < 0.f)
    t = t1;
  else if (t1 < 0.f && t2 > 0.f)
    t = t2;
  else
    return SphereHit{false, 0.f};
  return SphereHit{true, t};
}

// One 64-byte inner node as two 256-bit loads (LDG.E.256, sm_100+): divergent node fetches are
// bound by L1 wavefronts (one per lane and instruction), so halving the instruction count halves them.
#ifndef TUTU_NODE_LOAD_128
__device__ __forceinline__ void load_node(const float4* __restrict__ n, float4& a, float4& b, float4& c, int4& k) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(n));
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w), "=r"(k.x), "=r"(k.y), "=r"(k.z), "=r"(k.w)
               : "l"(n + 2));
}
#else
__device__ __forceinline__ void load_node(const float4* __restrict__ n, float4& a, float4& b, float4& c, int4& k) {
  a = __ldg(n + 0);
  b = __ldg(n + 1);
  c = __ldg(n + 2);
  k = __ldg(reinterpret_cast<const int4*>(n + 3));
}
#endif

struct VisitCount {
  unsigned nodes = 0, prims = 0;
};

// Slab test for "regular" rays (every 1/d finite): then no NaN can appear, the reference's
// swap-if-negative equals min/max of the two plane distances and its ternary max/min chains equal
// fmaxf/fminf (up to the sign of a zero, which no comparison below distinguishes), so the same
// decision is reached with FMNMX instead of FSETP+FSEL pairs.  Same subtract-multiply roundings.
__device__ __forceinline__ bool box_test_regular(const RayPre& p, float lox, float loy, float loz,
                                                 float hix, float hiy, float hiz, float& t_enter) {
  const float ax = __fmul_rn(__fsub_rn(lox, p.ox), p.ix), bx = __fmul_rn(__fsub_rn(hix, p.ox), p.ix);
  const float ay = __fmul_rn(__fsub_rn(loy, p.oy), p.iy), by = __fmul_rn(__fsub_rn(hiy, p.oy), p.iy);
  const float az = __fmul_rn(__fsub_rn(loz, p.oz), p.iz), bz = __fmul_rn(__fsub_rn(hiz, p.oz), p.iz);
  t_enter = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
  const float t_exit = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
  return t_enter <= t_exit && t_exit >= 0.f;
}

__device__ __forceinline__ bool ray_is_regular(const RayPre& p) {
  return isfinite(p.ix) && isfinite(p.iy) && isfinite(p.iz);
}

template <bool REGULAR>
__device__ __forceinline__ bool box_any(const RayPre& p, float lox, float loy, float loz, float hix,
                                        float hiy, float hiz, float& t_enter) {
  if (REGULAR) return box_test_regular(p, lox, loy, loz, hix, hiy, hiz, t_enter);
  return box_test(p, lox, loy, loz, hix, hiy, hiz, t_enter);
}

// Slab test for regular rays with Blackwell's packed fp32 instructions (FADD2 / FMUL2, sm_100+):
// both planes of an axis go through one subtract and one multiply.  add.rn.f32x2(p, -o) is the IEEE
// p - o and mul.rn.f32x2 the IEEE product, so the two t values per axis are bit-identical to
// box_test_regular's; only the issue-slot count halves (these kernels are issue bound).
struct RayPre2 {
  float2 nox, noy, noz;  // (-o, -o)
  float2 ix, iy, iz;     // (1/d, 1/d)
};
__device__ __forceinline__ RayPre2 make_pre2(const RayPre& p) {
  RayPre2 q;
  q.nox = make_float2(-p.ox, -p.ox), q.noy = make_float2(-p.oy, -p.oy), q.noz = make_float2(-p.oz, -p.oz);
  q.ix = make_float2(p.ix, p.ix), q.iy = make_float2(p.iy, p.iy), q.iz = make_float2(p.iz, p.iz);
  return q;
}
__device__ __forceinline__ bool box_test_packed(const RayPre2& q, float2 bx, float2 by, float2 bz, float& t_enter) {
  const float2 tx = __fmul2_rn(__fadd2_rn(bx, q.nox), q.ix);
  const float2 ty = __fmul2_rn(__fadd2_rn(by, q.noy), q.iy);
  const float2 tz = __fmul2_rn(__fadd2_rn(bz, q.noz), q.iz);
  t_enter = fmaxf(fminf(tx.x, tx.y), fmaxf(fminf(ty.x, ty.y), fminf(tz.x, tz.y)));
  const float t_exit = fminf(fmaxf(tx.x, tx.y), fminf(fmaxf(ty.x, ty.y), fmaxf(tz.x, tz.y)));
  return t_enter <= t_exit && t_exit >= 0.f;
}

// Both child boxes of a 64-byte inner node.  Node layout (host_scene.cpp: set_node_boxes): per child and
// axis {lo, hi} pairs, so that a regular ray tests an axis with one FADD2 and one FMUL2:
//   a = {l.x.lo, l.x.hi, l.y.lo, l.y.hi}  b = {l.z.lo, l.z.hi, r.x.lo, r.x.hi}  c = {r.y.lo, r.y.hi, r.z.lo, r.z.hi}
template <bool REGULAR>
__device__ __forceinline__ void node_boxes(const RayPre& p, const float4 a, const float4 b, const float4 c, bool& hl,
                                           float& tl, bool& hr, float& tr) {
  if (REGULAR) {
    const RayPre2 q = make_pre2(p);
    hl = box_test_packed(q, make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(b.x, b.y), tl);
    hr = box_test_packed(q, make_float2(b.z, b.w), make_float2(c.x, c.y), make_float2(c.z, c.w), tr);
  } else {
    hl = box_test(p, a.x, a.z, b.x, a.y, a.w, b.y, tl);
    hr = box_test(p, b.z, c.x, c.z, b.w, c.y, c.w, tr);
  }
}

// Per-ray traversal state.  The walk is cut into "rounds" (descend inner nodes until a leaf is
// reached, test that one leaf, pop) so that a warp can (a) run the inner-node code and the
// primitive code as two convergent phases instead of interleaving them lane by lane, and (b) hand
// a finished lane a new ray between rounds (persistent threads, see trace_persistent()).
struct Walk {
  const float4* nodes;  // the node array this walk descends (DevScene::inner_fast for regular rays)
  Ray r;
  RayPre p;
  Hit best;
  float dis;      // any-hit distance limit
  int cur;
  int sp;
  bool regular;
};

// MODE 0: ordered (near child first) + pruned by the best t so far.  The reference never prunes, so
//         culling must not remove a primitive it would have reported.  In exact arithmetic a
//         triangle inside a box is hit no earlier than the box is entered; in fp32 the slab
//         distances carry ~3 ulp of relative error, but a Cramer-rule t carries up to
//         ~10 eps / (sin(E1,E2) * |d.n|) (|d.n| >= 1e-4 is enforced by the parallel test), i.e.
//         up to a fraction of a percent for extreme grazing rays.  A subtree is therefore skipped
//         only when t_enter > best_t * (1 + prune_rel) + prune_abs (defaults 2^-10 and 2^-9 of
//         the longest triangle edge); equal t is resolved by the lower DFS leaf slot, which is
//         the reference's tie rule.  Measured on the 2 x 2^24 synthetic rays (profiles/): slack 0
//         -> 16 rays differ from the literal walk (shared-edge hits with equal / 1-ulp-apart t),
//         slack 2^-10 -> 0 differ at +4 % node visits, 2^-7 -> 0 differ at +25 %.  tests/ and
//         bench.py compare MODE 0 with MODE 1 bit for bit on the full-size batches.
// MODE 1: literal mirror of the reference recursion: left then right, nothing pruned.
// ANY:    hasIntersection — first accepted leaf ends the walk (the boolean is order independent).
template <bool ANY>
__device__ __forceinline__ float prune_limit(const DevScene& sc, const Walk& w) {
  const float b = ANY ? w.dis : w.best.t;
  return fmaf(b, sc.prune_rel, b + sc.prune_abs);
}

// Starts a walk: false when the ray cannot hit anything (empty scene / root box missed).
// use_fast = false keeps the walk on the reference topology (literal mode).
__device__ __forceinline__ bool walk_begin(const DevScene& sc, Walk& w, const Ray& r, float dis, bool use_fast = true) {
  w.r = r;
  w.dis = dis;
  w.best.t = FLT_MAX;
  w.best.u = 0.f;
  w.best.v = 0.f;
  w.best.slot = -1;
  w.sp = 0;
  w.cur = sc.root_ref;
  w.nodes = sc.inner;
  if (sc.empty) return false;
  w.p = make_pre(r);
  w.regular = ray_is_regular(w.p);
  if (use_fast && w.regular) {
    w.cur = sc.root_ref_fast;
    w.nodes = sc.inner_fast;
  }
  float te;
  return box_test(w.p, sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1],
                  sc.root_hi[2], te);
}

template <bool ANY, int MODE>
__device__ __forceinline__ bool walk_pop(const DevScene& sc, Walk& w, const int* stack_ref,
                                         const float* stack_t) {
  for (;;) {
    if (w.sp == 0) return false;
    --w.sp;
    w.cur = stack_ref[w.sp];
    if (MODE == 0 && !ANY && stack_t[w.sp] > prune_limit<ANY>(sc, w)) continue;  // entered after the best hit
    return true;
  }
}

// One round.  Returns false when the walk is over (result in w.best; for ANY best.slot >= 0 means
// "blocked").  REGULAR selects the slab-test flavour for the whole warp-round.
template <bool ANY, int MODE, bool COUNT, bool REGULAR>
__device__ __forceinline__ bool walk_round(const DevScene& sc, Walk& w, int* stack_ref, float* stack_t,
                                           VisitCount* vc) {
  // phase 1: inner nodes until this lane holds a leaf
  while (w.cur >= 0) {
    const float4* n = w.nodes + 4 * (size_t)w.cur;
    float4 a, b, c;
    int4 k;
    load_node(n, a, b, c, k);
    if (COUNT) vc->nodes++;
    float tl, tr;
    bool hl, hr;
    node_boxes<REGULAR>(w.p, a, b, c, hl, tl, hr, tr);
    if (MODE == 0) {
      const float lim = prune_limit<ANY>(sc, w);
      hl = hl && !(tl > lim);
      hr = hr && !(tr > lim);
    }
    if (hl && hr) {
      const bool left_first = (MODE == 1) || !(tr < tl);
      stack_ref[w.sp] = left_first ? k.y : k.x;
      stack_t[w.sp] = left_first ? tr : tl;
      ++w.sp;
      w.cur = left_first ? k.x : k.y;
    } else if (hl) {
      w.cur = k.x;
    } else if (hr) {
      w.cur = k.y;
    } else if (!walk_pop<ANY, MODE>(sc, w, stack_ref, stack_t)) {
      return false;
    }
  }
  // phase 2: one primitive
  {
    const uint32_t code = ~(uint32_t)w.cur;
    const uint32_t slot = code & kSlotMask;
    const float4* g = sc.geom + 3 * (size_t)slot;
    if (COUNT) vc->prims++;
    float t, u = 0.f, v = 0.f;
    bool hit;
    if (code & kSphereBit) {
      const SphereHit sh = sphere_test(g, w.r);
      hit = sh.hit;
      t = sh.t;
    } else {
      hit = tri_test(g, w.r, t, u, v);
    }
    if (hit) {
      if (ANY) {
        // inter.t < dis && !FLOAT_EQUAL(inter.t, dis)
        if (t < w.dis && !(fabsf(__fsub_rn(t, w.dis)) < 0.0001f)) {
          w.best.t = t;
          w.best.slot = (int)code;
          return false;
        }
      } else if (t < w.best.t || (t == w.best.t && (int)slot < (w.best.slot & (int)kSlotMask))) {
        // linter.t <= rinter.t ? linter : rinter  ==  lowest DFS leaf among equal t
        w.best.t = t;
        w.best.u = u;
        w.best.v = v;
        w.best.slot = (int)code;
      }
    }
  }
  return walk_pop<ANY, MODE>(sc, w, stack_ref, stack_t);
}

// Whole walk of one ray on one thread (used for the rare inline shadow ray of the shade kernel,
// the literal MODE 1 kernels and the visit counters).
template <bool ANY, int MODE, bool COUNT>
__device__ __forceinline__ bool traverse(const DevScene& sc, const Ray& r, float dis, Hit& best,
                                         VisitCount* vc) {
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  if (walk_begin(sc, w, r, dis, MODE == 0)) {
    if (w.regular) {
      while (walk_round<ANY, MODE, COUNT, true>(sc, w, stack_ref, stack_t, vc)) {
      }
    } else {
      while (walk_round<ANY, MODE, COUNT, false>(sc, w, stack_ref, stack_t, vc)) {
      }
    }
  }
  best = w.best;
  return best.slot >= 0;
}

// ---- single-thread walk flavours (compared in tests/tools/gpu_variants.py) ---------------------------
// LOOP: one loop whose body handles either an inner node or a leaf (per lane).
template <bool ANY, int MODE, bool REGULAR>
__device__ __forceinline__ void walk_loop(const DevScene& sc, Walk& w, int* stack_ref, float* stack_t) {
  for (;;) {
    if (w.cur >= 0) {
      const float4* n = w.nodes + 4 * (size_t)w.cur;
      float4 a, b, c;
      int4 k;
      load_node(n, a, b, c, k);
      float tl, tr;
      bool hl, hr;
      node_boxes<REGULAR>(w.p, a, b, c, hl, tl, hr, tr);
      if (MODE == 0) {
        const float lim = prune_limit<ANY>(sc, w);
        hl = hl && !(tl > lim);
        hr = hr && !(tr > lim);
      }
      if (hl && hr) {
        const bool left_first = (MODE == 1) || !(tr < tl);
        stack_ref[w.sp] = left_first ? k.y : k.x;
        stack_t[w.sp] = left_first ? tr : tl;
        ++w.sp;
        w.cur = left_first ? k.x : k.y;
        continue;
      }
      if (hl) {
        w.cur = k.x;
        continue;
      }
      if (hr) {
        w.cur = k.y;
        continue;
      }
    } else {
      const uint32_t code = ~(uint32_t)w.cur;
      const uint32_t slot = code & kSlotMask;
      const float4* g = sc.geom + 3 * (size_t)slot;
      float t, u = 0.f, v = 0.f;
      bool hit;
      if (code & kSphereBit) {
        const SphereHit sh = sphere_test(g, w.r);
        hit = sh.hit;
        t = sh.t;
      } else {
        hit = tri_test(g, w.r, t, u, v);
      }
      if (hit) {
        if (ANY) {
          if (t < w.dis && !(fabsf(__fsub_rn(t, w.dis)) < 0.0001f)) {
            w.best.t = t;
            w.best.slot = (int)code;
            return;
          }
        } else if (t < w.best.t || (t == w.best.t && (int)slot < (w.best.slot & (int)kSlotMask))) {
          w.best.t = t;
          w.best.u = u;
          w.best.v = v;
          w.best.slot = (int)code;
        }
      }
    }
    if (!walk_pop<ANY, MODE>(sc, w, stack_ref, stack_t)) return;
  }
}


// ---- structured walks --------------------------------------------------------------------------
// ncu source view of walk_loop (profiles/r01_closest_lanes.txt): the node code runs with 16 of 32
// lanes, but its four exits (`continue` out of nested ifs) reconverge only at the loop head, so
// the stack pop runs once per exit path with 3-4 lanes, the push with 5, and a leaf test whenever a
// single lane holds a leaf: 10 lanes per instruction on average.  Below, every iteration is
// "one node or one leaf, then (for everybody who needs it) one pop", written as structured
// if/else so the compiler reconverges the warp before the pop.
template <bool ANY>
__device__ __forceinline__ bool leaf_step(const DevScene& sc, Walk& w, int leaf_ref) {
  const uint32_t code = ~(uint32_t)leaf_ref;
  const uint32_t slot = code & kSlotMask;
  const float4* g = sc.geom + 3 * (size_t)slot;
  float t, u = 0.f, v = 0.f;
  bool hit;
  if (code & kSphereBit) {
    const SphereHit sh = sphere_test(g, w.r);
    hit = sh.hit;
    t = sh.t;
  } else {
    hit = tri_test(g, w.r, t, u, v);
  }
  if (hit) {
    if (ANY) {
      if (t < w.dis && !(fabsf(__fsub_rn(t, w.dis)) < 0.0001f)) {
        w.best.t = t;
        w.best.slot = (int)code;
        return true;  // blocked: the walk is over
      }
    } else if (t < w.best.t || (t == w.best.t && (int)slot < (w.best.slot & (int)kSlotMask))) {
      w.best.t = t;
      w.best.u = u;
      w.best.v = v;
      w.best.slot = (int)code;
    }
  }
  return false;
}

// node step: returns true when the lane has to pop
template <bool ANY, bool REGULAR>
__device__ __forceinline__ bool node_step(const DevScene& sc, Walk& w, int* stack_ref, float* stack_t) {
  const float4* n = w.nodes + 4 * (size_t)w.cur;
  float4 a, b, c;
  int4 k;
  load_node(n, a, b, c, k);
  float tl, tr;
  bool hl, hr;
  node_boxes<REGULAR>(w.p, a, b, c, hl, tl, hr, tr);
  const float lim = prune_limit<ANY>(sc, w);
  hl = hl && !(tl > lim);
  hr = hr && !(tr > lim);
  const bool swap = hr && (!hl || tr < tl);  // near child = right
  const int near = swap ? k.y : k.x, far = swap ? k.x : k.y;
  if (hl && hr) {
    stack_ref[w.sp] = far;
    stack_t[w.sp] = swap ? tl : tr;
    ++w.sp;
  }
  if (hl || hr) {
    w.cur = near;
    return false;
  }
  return true;
}

template <bool ANY, bool REGULAR>
__device__ __forceinline__ void walk_structured(const DevScene& sc, Walk& w, int* stack_ref, float* stack_t) {
  for (;;) {
    bool need_pop;
    if (w.cur >= 0) {
      need_pop = node_step<ANY, REGULAR>(sc, w, stack_ref, stack_t);
    } else {
      if (leaf_step<ANY>(sc, w, w.cur)) return;
      need_pop = true;
    }
    if (need_pop && !walk_pop<ANY, 0>(sc, w, stack_ref, stack_t)) return;
  }
}

// ---- traversal stack in shared memory ------------------------------------------------------------
// ncu (profiles/r01_rays_full.txt): the per-thread local-memory stack costs 0.44e9 of the 2.3e9 L1
// tag-stage wavefronts of the closest-hit kernel and 41 GB of L2 sectors (3 useful bytes per
// 32-byte sector: every lane's 4-byte slot sits in its own sector once the lanes' stack depths
// differ).  In shared memory entry e of thread t lives at word (e * blockDim + t): any mix of
// depths is bank-conflict free, and {ref, t_enter} travel as one 64-bit word.
// The any-hit walk never prunes on pop (its limit is the constant `dis`, already applied when the
// entry was pushed), so its entries are the 32-bit refs alone.
template <bool ANY>
struct SharedStack;
template <>
struct SharedStack<false> {
  using Word = unsigned long long;
  Word* base;  // this thread's column: base[e * stride]
  unsigned stride;
  int sp = 0;
  __device__ __forceinline__ void push(int ref, float t) {
    base[(unsigned)sp * stride] = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)ref;
    ++sp;
  }
  __device__ __forceinline__ void pop(int& ref, float& t) {
    --sp;
    const unsigned long long v = base[(unsigned)sp * stride];
    ref = (int)(unsigned)v;
    t = __uint_as_float((unsigned)(v >> 32));
  }
};
template <>
struct SharedStack<true> {
  using Word = unsigned;
  Word* base;
  unsigned stride;
  int sp = 0;
  __device__ __forceinline__ void push(int ref, float) {
    base[(unsigned)sp * stride] = (unsigned)ref;
    ++sp;
  }
  __device__ __forceinline__ void pop(int& ref, float& t) {
    --sp;
    ref = (int)base[(unsigned)sp * stride];
    t = 0.f;
  }
};

// node part of a step with the shared-memory stack: true = the lane has to pop
template <bool ANY, bool REGULAR, class Stack>
__device__ __forceinline__ bool node_step_shared(const DevScene& sc, Walk& w, Stack& st) {
  const float4* n = w.nodes + 4 * (size_t)w.cur;
  float4 a, b, c;
  int4 k;
  load_node(n, a, b, c, k);
  float tl, tr;
  bool hl, hr;
  node_boxes<REGULAR>(w.p, a, b, c, hl, tl, hr, tr);
  const float lim = prune_limit<ANY>(sc, w);
  hl = hl && !(tl > lim);
  hr = hr && !(tr > lim);
  const bool swap = hr && (!hl || tr < tl);
  if (hl && hr) st.push(swap ? k.x : k.y, swap ? tl : tr);
  if (hl || hr) {
    w.cur = swap ? k.y : k.x;
    return false;
  }
  return true;
}

template <bool ANY, bool REGULAR>
__device__ __forceinline__ void walk_shared(const DevScene& sc, Walk& w, SharedStack<ANY>& st) {
  for (;;) {
    bool need_pop;
    if (w.cur >= 0) {
      need_pop = node_step_shared<ANY, REGULAR>(sc, w, st);
    } else {
      if (leaf_step<ANY>(sc, w, w.cur)) return;
      need_pop = true;
    }
    if (need_pop) {
      for (;;) {
        if (st.sp == 0) return;
        float t;
        st.pop(w.cur, t);
        if (!ANY && t > prune_limit<ANY>(sc, w)) continue;
        break;
      }
    }
  }
}

// whole walk of one ray with the shared-memory stack.  smem: blockDim * (tree depth + 1) words of
// SharedStack<ANY>::Word (a ray pushes at most one entry per tree level).
template <bool ANY>
__device__ __forceinline__ bool traverse_shared(const DevScene& sc, const Ray& r, float dis, Hit& best, void* smem) {
  Walk w;
  SharedStack<ANY> st;
  st.base = reinterpret_cast<typename SharedStack<ANY>::Word*>(smem) + threadIdx.x;
  st.stride = blockDim.x;
  if (walk_begin(sc, w, r, dis)) {
    if (w.regular)
      walk_shared<ANY, true>(sc, w, st);
    else
      walk_shared<ANY, false>(sc, w, st);
  }
  best = w.best;
  return best.slot >= 0;
}

// ---- warp walk with batched leaf tests ---------------------------------------------------------
// ncu source view (profiles/r01_bdpt_lanes.txt, Veach room): the 85-instruction primitive test runs
// with 2 of 32 lanes — in any one iteration few lanes stand on a leaf — and costs 36 % of the
// issue slots; the node code runs with 12.  Here the lanes of a warp walk together: a lane that
// reaches a leaf WAITS (it would idle through the others' node steps anyway) until at least
// kLeafBatch lanes wait or nobody is left on an inner node; then all waiting lanes test their
// primitives at once.  Waiting loses no pruning information, so no extra nodes are visited.
// All 32 lanes call this together; `valid` = the lane carries a ray.
constexpr int kLeafBatch = 8;
template <bool ANY>
__device__ __forceinline__ bool traverse_batched(const DevScene& sc, const Ray& r, float dis, bool valid, Hit& best,
                                                 void* smem) {
  Walk w;
  SharedStack<ANY> st;
  st.base = reinterpret_cast<typename SharedStack<ANY>::Word*>(smem) + threadIdx.x;
  st.stride = blockDim.x;
  w.best.t = FLT_MAX, w.best.u = 0.f, w.best.v = 0.f, w.best.slot = -1;
  w.regular = true;
  w.cur = 0;
  bool done = !valid || !walk_begin(sc, w, r, dis);
  const bool all_regular = __all_sync(0xFFFFFFFFu, done || w.regular);
  for (;;) {
    const unsigned at_node = __ballot_sync(0xFFFFFFFFu, !done && w.cur >= 0);
    const unsigned at_leaf = __ballot_sync(0xFFFFFFFFu, !done && w.cur < 0);
    if ((at_node | at_leaf) == 0u) break;
    const int n_leaf = __popc(at_leaf), n_node = __popc(at_node);
    bool need_pop = false;
    if (n_leaf >= sc.leaf_batch || n_leaf >= n_node) {  // enough lanes wait (or nobody walks): test the leaves
      if (!done && w.cur < 0) {
        if (leaf_step<ANY>(sc, w, w.cur)) done = true;
        need_pop = !done;
      }
    } else if (!done && w.cur >= 0) {
      need_pop = all_regular ? node_step_shared<ANY, true>(sc, w, st) : node_step_shared<ANY, false>(sc, w, st);
    }
    if (need_pop) {
      for (;;) {
        if (st.sp == 0) {
          done = true;
          break;
        }
        float t;
        st.pop(w.cur, t);
        if (!ANY && t > prune_limit<ANY>(sc, w)) continue;
        break;
      }
    }
  }
  best = w.best;
  return best.slot >= 0;
}

// ---- traversal stack in local memory (same interface as SharedStack) -------------------------------
template <bool ANY>
struct LocalStack;
template <>
struct LocalStack<false> {
  unsigned long long e[kStackSize];
  int sp = 0;
  __device__ __forceinline__ void push(int ref, float t) {
    e[sp] = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned)ref;
    ++sp;
  }
  __device__ __forceinline__ void pop(int& ref, float& t) {
    --sp;
    const unsigned long long v = e[sp];
    ref = (int)(unsigned)v;
    t = __uint_as_float((unsigned)(v >> 32));
  }
};
template <>
struct LocalStack<true> {
  unsigned e[kStackSize];
  int sp = 0;
  __device__ __forceinline__ void push(int ref, float) {
    e[sp] = (unsigned)ref;
    ++sp;
  }
  __device__ __forceinline__ void pop(int& ref, float& t) {
    --sp;
    ref = (int)e[sp];
    t = 0.f;
  }
};

// ---- queue tracer: persistent lanes, phase-separated steps ------------------------------------------
// ncu on the incoherent queues (profiles/r02_q_extend_lanes.txt: 5.8 of 32 lanes; r02_glass_extend_lanes.txt: 9): a
// warp that walks 32 rays to completion idles behind its longest walk, and within a step its lanes are spread over
// node code, primitive test and pop.  Each remedy alone lost its measurement: handing idle lanes
// new rays desynchronises the warp, so that EVERY step runs all three code paths for a few lanes each
// (trace_variants.cuh: trace_refill); making lanes wait on their leaf until enough wait leaves them idle next to the
// finished ones (traverse_batched on queues).  Together they fit: a warp owns a chunk of the queue and refills its idle
// lanes from it whenever `refill_min` are idle, and every iteration of the warp is ONE phase — a node step for the
// lanes that stand on an inner node, or, once `leaf_batch` lanes wait on a leaf (or nobody stands on a node), the
// primitive tests of the waiting lanes — followed by the pops of the lanes that need one.
//   load(i, ray, dis) -> false for a dead queue entry (the callee has then written its result itself)
//   store(i, hit, any)   any = best.slot >= 0
constexpr unsigned kQueueChunk = 256;
constexpr int kLanesLeafBatch = 16;  // Veach room, Msamples/s at 8 / 16 / 20: 73.2 / 74.1 / 71.9
constexpr int kLanesRefillMin = 8;
template <bool ANY, class Stack, class Load, class Store>
__device__ __forceinline__ void trace_queue_lanes(const DevScene& sc, unsigned n, unsigned long long* cursor, Stack& st,
                                                  Load load, Store store) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  Walk w;
  w.cur = 0;
  w.regular = true;
  bool active = false;
  unsigned my = 0u;
  unsigned chunk_next = 0u, chunk_end = 0u;  // warp-uniform
  bool exhausted = false;                    // warp-uniform
  bool all_regular = true;                   // warp-uniform; re-evaluated at every refill (stale "false" is only slower)
  for (;;) {
    const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (!exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= sc.refill_min)) {
      if (chunk_next == chunk_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kQueueChunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        chunk_next = base < n ? (unsigned)base : n;
        chunk_end = base + kQueueChunk < n ? (unsigned)base + kQueueChunk : n;
        if (chunk_next >= n) exhausted = true;
      }
      const unsigned avail = chunk_end - chunk_next < 32u ? chunk_end - chunk_next : 32u;
      const unsigned rank = (unsigned)__popc(idle & lt);
      if (!active && rank < avail) {
        my = chunk_next + rank;
        Ray r;
        float dis = 0.f;
        if (load(my, r, dis)) {
          st.sp = 0;
          if (walk_begin(sc, w, r, dis))
            active = true;
          else
            store(my, w.best, false);  // missed the scene box: w.best is the miss record
        }
      }
      const unsigned want = (unsigned)__popc(idle);
      chunk_next += want < avail ? want : avail;
      all_regular = __all_sync(0xFFFFFFFFu, !active || w.regular);
    }
    const unsigned at_node = __ballot_sync(0xFFFFFFFFu, active && w.cur >= 0);
    const unsigned at_leaf = __ballot_sync(0xFFFFFFFFu, active && w.cur < 0);
    if ((at_node | at_leaf) == 0u) {
      if (exhausted) return;
      continue;
    }
    bool need_pop = false;
    if (__popc(at_leaf) >= sc.lanes_leaf_batch || at_node == 0u) {
      if (active && w.cur < 0) {
        if (leaf_step<ANY>(sc, w, w.cur)) {  // any-hit: blocked
          active = false;
          store(my, w.best, true);
        } else {
          need_pop = true;
        }
      }
    } else if (active && w.cur >= 0) {
      need_pop = all_regular ? node_step_shared<ANY, true>(sc, w, st) : node_step_shared<ANY, false>(sc, w, st);
    }
    if (need_pop) {
      for (;;) {
        if (st.sp == 0) {
          active = false;
          store(my, w.best, w.best.slot >= 0);
          break;
        }
        float t;
        st.pop(w.cur, t);
        if (!ANY && t > prune_limit<ANY>(sc, w)) continue;
        break;
      }
    }
  }
}

// Whole walk of one ray with a local-memory stack and the exact (NaN-literal) slab test: the walk of irregular rays
// on scenes that otherwise use the flat small-scene tests.
template <bool ANY>
__device__ __forceinline__ bool traverse_exact(const DevScene& sc, const Ray& r, float dis, Hit& best) {
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  if (walk_begin(sc, w, r, dis)) walk_loop<ANY, 0, false>(sc, w, stack_ref, stack_t);
  best = w.best;
  return best.slot >= 0;
}
// Structured walk (one node or leaf, then one pop; local-memory stack): the any-hit walk of the queue kernels.
template <bool ANY>
__device__ __forceinline__ bool traverse_structured(const DevScene& sc, const Ray& r, float dis, Hit& best) {
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  if (walk_begin(sc, w, r, dis)) {
    if (w.regular)
      walk_structured<ANY, true>(sc, w, stack_ref, stack_t);
    else
      walk_structured<ANY, false>(sc, w, stack_ref, stack_t);
  }
  best = w.best;
  return best.slot >= 0;
}

// ---- small scenes (<= 32 primitives): no tree walk at all --------------------------------------
// For a regular ray (every 1/d finite, so no NaN anywhere in the slab arithmetic) the slab test
// is monotone in the box: fl((plane - o) * inv) is non-decreasing in `plane`, every inner box of
// the reference tree is the exact fmin/fmax union of its children (BVH.hpp:65,119), hence
// t_enter(parent) <= t_enter(child) <= t_exit(child) <= t_exit(parent): a hit leaf box implies that
// every ancestor box is hit.  The reference's answer is therefore simply the closest primitive
// among those whose OWN leaf box and primitive test both pass (ties: lowest DFS slot).  With 32
// leaves that is at most 32 slab tests (one per distinct box) against constants read straight from the kernel-parameter bank
// (fully convergent, no loads, no stack) followed by the primitive tests of the few set bits.
// Irregular rays take the literal tree walk.
constexpr int kSmallMax = 32;
struct SmallScene {
  int n;        // primitives; 0 = not usable (scene too large or empty)
  int n_boxes;  // distinct leaf boxes (bitwise): the two triangles of a quad have the same bounds, so the
                // Cornell box's 32 leaves are 16 slab tests
  float2 box[kSmallMax][3];  // distinct leaf boxes, per axis {lo, hi} (a packed-fp32 operand)
  unsigned slots[kSmallMax]; // DFS slots of the primitives that own box k (0 for the unused tail)
};

// one group of 4 distinct boxes, indices known at compile time (operands straight from the constant bank)
template <bool ANY, int K0>
__device__ __forceinline__ unsigned small_box_group(const SmallScene& ss, const RayPre2& q, float lim) {
  unsigned mask = 0u;
#pragma unroll
  for (int k = K0; k < K0 + 4; ++k) {
    // box_test_packed with the two comparisons folded: t_enter <= t_exit && t_exit >= 0  <=>
    // max(t_enter, 0) <= t_exit (no NaN for a regular ray); any-hit: ... && t_enter <= lim
    const float2 tx = __fmul2_rn(__fadd2_rn(ss.box[k][0], q.nox), q.ix);
    const float2 ty = __fmul2_rn(__fadd2_rn(ss.box[k][1], q.noy), q.iy);
    const float2 tz = __fmul2_rn(__fadd2_rn(ss.box[k][2], q.noz), q.iz);
    const float te = fmaxf(fminf(tx.x, tx.y), fmaxf(fminf(ty.x, ty.y), fminf(tz.x, tz.y)));
    float t_exit = fminf(fmaxf(tx.x, tx.y), fminf(fmaxf(ty.x, ty.y), fmaxf(tz.x, tz.y)));
    if (ANY) t_exit = te <= lim ? t_exit : -1.f;  // beyond the limit: cannot hold a blocker
    if (fmaxf(te, 0.f) <= t_exit) mask |= ss.slots[k];
  }
  return mask;
}

// Candidate mask of a regular ray: DFS slots of the primitives whose leaf box the ray hits.
template <bool ANY>
__device__ __forceinline__ unsigned small_candidates(const DevScene& sc, const SmallScene& ss, const RayPre& p, float dis) {
  const RayPre2 q = make_pre2(p);
  const float lim = ANY ? fmaf(dis, sc.prune_rel, dis + sc.prune_abs) : 0.f;
  unsigned mask = small_box_group<ANY, 0>(ss, q, lim);  // n_boxes is uniform: whole groups are skipped
  if (ss.n_boxes > 4) mask |= small_box_group<ANY, 4>(ss, q, lim);
  if (ss.n_boxes > 8) mask |= small_box_group<ANY, 8>(ss, q, lim);
  if (ss.n_boxes > 12) mask |= small_box_group<ANY, 12>(ss, q, lim);
  if (ss.n_boxes > 16) mask |= small_box_group<ANY, 16>(ss, q, lim);
  if (ss.n_boxes > 20) mask |= small_box_group<ANY, 20>(ss, q, lim);
  if (ss.n_boxes > 24) mask |= small_box_group<ANY, 24>(ss, q, lim);
  if (ss.n_boxes > 28) mask |= small_box_group<ANY, 28>(ss, q, lim);
  return mask;
}

// Primitive test of the lowest candidate.  Candidates are taken in increasing slot order: strict '<'
// keeps the lowest slot on ties.  Returns true when an any-hit ray is blocked (mask is then cleared).
template <bool ANY>
__device__ __forceinline__ bool small_test_next(const DevScene& sc, const Ray& r, float dis, unsigned& mask, Hit& best) {
  const int slot = __ffs(mask) - 1;
  mask &= mask - 1u;
  const float4* g = sc.geom + 3 * slot;
  const bool sphere = (sc.sphere_mask >> slot) & 1u;
  float t, u = 0.f, v = 0.f;
  bool hit;
  if (sphere) {
    const SphereHit sh = sphere_test(g, r);
    hit = sh.hit;
    t = sh.t;
  } else {
    hit = tri_test(g, r, t, u, v);
  }
  if (hit) {
    if (ANY) {
      if (t < dis && !(fabsf(__fsub_rn(t, dis)) < 0.0001f)) {
        best.t = t;
        best.slot = slot | (sphere ? (int)kSphereBit : 0);
        mask = 0u;
        return true;
      }
    } else if (t < best.t) {
      best.t = t;
      best.u = u;
      best.v = v;
      best.slot = slot | (sphere ? (int)kSphereBit : 0);
    }
  }
  return false;
}

template <bool ANY>
__device__ __forceinline__ bool traverse_small(const DevScene& sc, const SmallScene& ss, const Ray& r,
                                               float dis, Hit& best) {
  const RayPre p = make_pre(r);
  if (!ray_is_regular(p)) return traverse_exact<ANY>(sc, r, dis, best);
  best.t = FLT_MAX;
  best.u = 0.f;
  best.v = 0.f;
  best.slot = -1;
  unsigned mask = small_candidates<ANY>(sc, ss, p, dis);
  while (mask)
    if (small_test_next<ANY>(sc, r, dis, mask, best)) return true;
  return best.slot >= 0;
}

// Two-phase form for the queue kernels (wavefront.cuh).  ncu on the Cornell box (profiles/r01e_resident_*):
// the primitive loop above runs 8.8 rounds per warp with 10 of 32 lanes — 88 % of the rays have the two
// candidates of one wall quad, but nearly every warp holds a ray that crosses a block's bounds and has
// 6-10.  So a warp tests at most kSmallFirst candidates per ray in its first pass, parks the rays that
// still have candidates (state in shared memory, warp-private, no atomics) and finishes 32 parked rays
// at a time with all lanes busy.
#ifndef TUTU_SMALL_FIRST
#define TUTU_SMALL_FIRST 2
#endif
#ifndef TUTU_SMALL_NEXT
#define TUTU_SMALL_NEXT 2
#endif
constexpr int kSmallFirst = TUTU_SMALL_FIRST;
constexpr unsigned kParkCap = 64;  // a warp parks at most 31 + 32 rays before it drains 32
struct SmallPark {                 // per warp
  float4 o[kParkCap];              // o.xyz, dis
  float4 d[kParkCap];              // d.xyz, bits(queue index)
  float4 best[kParkCap];           // t, u, v, bits(slot)
  unsigned mask[kParkCap];
};

// begin: candidate mask + first pass.  Returns true when the ray is finished (result in best / blocked).
template <bool ANY>
__device__ __forceinline__ bool small_first_pass(const DevScene& sc, const SmallScene& ss, const Ray& r, float dis,
                                                 Hit& best, unsigned& mask, bool& blocked) {
  const RayPre p = make_pre(r);
  mask = 0u;
  if (!ray_is_regular(p)) {
    blocked = traverse_exact<ANY>(sc, r, dis, best);
    return true;
  }
  best.t = FLT_MAX;
  best.u = 0.f;
  best.v = 0.f;
  best.slot = -1;
  blocked = false;
  mask = small_candidates<ANY>(sc, ss, p, dis);
#pragma unroll
  for (int k = 0; k < kSmallFirst; ++k)
    if (mask && small_test_next<ANY>(sc, r, dis, mask, best)) blocked = true;
  return mask == 0u;
}

// Warp-collective: lanes with `more` append their ray to the warp's park; returns the new count.
__device__ __forceinline__ unsigned small_park_push(SmallPark& pk, unsigned cnt, bool more, const Ray& r, float dis,
                                                    unsigned index, const Hit& best, unsigned mask) {
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, more);
  if (more) {
    const unsigned pos = cnt + (unsigned)__popc(bal & ((1u << (threadIdx.x & 31u)) - 1u));
    pk.o[pos] = make_float4(r.ox, r.oy, r.oz, dis);
    pk.d[pos] = make_float4(r.dx, r.dy, r.dz, __uint_as_float(index));
    pk.best[pos] = make_float4(best.t, best.u, best.v, __int_as_float(best.slot));
    pk.mask[pos] = mask;
  }
  __syncwarp();
  return cnt + (unsigned)__popc(bal);
}

// Warp-collective: the last min(cnt, 32) parked rays get kSmallNext more tests each (all their remaining
// candidates when `finish`); finished rays go to sink(index, ray, dis, best, blocked), the others are
// parked again.  Returns the new count.
constexpr int kSmallNext = TUTU_SMALL_NEXT;
template <bool ANY, class Sink>
__device__ __forceinline__ unsigned small_park_drain(const DevScene& sc, SmallPark& pk, unsigned cnt, bool finish,
                                                     Sink sink) {
  const unsigned take = cnt < 32u ? cnt : 32u;
  const unsigned lane = threadIdx.x & 31u;
  const bool mine = lane < take;
  Ray r{};
  Hit best{};
  unsigned mask = 0u, index = 0u;
  float dis = 0.f;
  if (mine) {
    const unsigned e = cnt - take + lane;
    const float4 o = pk.o[e], d = pk.d[e], b = pk.best[e];
    mask = pk.mask[e];
    r = Ray{o.x, o.y, o.z, d.x, d.y, d.z};
    dis = o.w;
    index = __float_as_uint(d.w);
    best = Hit{b.x, b.y, b.z, __float_as_int(b.w)};
    bool blocked = false;
    if (finish) {
      while (mask)
        if (small_test_next<ANY>(sc, r, dis, mask, best)) blocked = true;
    } else {
#pragma unroll
      for (int k = 0; k < kSmallNext; ++k)
        if (mask && small_test_next<ANY>(sc, r, dis, mask, best)) blocked = true;
    }
    if (mask == 0u) sink(index, r, dis, best, ANY ? blocked : best.slot >= 0);
  }
  __syncwarp();  // every lane has read its entry before the survivors are written back
  return small_park_push(pk, cnt - take, mine && mask != 0u, r, dis, index, best, mask);
}

}  // namespace tutu

#include "wide.cuh"
#ifdef TUTU_EXPERIMENTS
#include "trace_variants.cuh"  // walk flavours that lost their measurements; not in the shipped library
#endif
