// wide.cuh — walk of the compressed 8-wide traversal tree (tutu_internal.hpp: WideNode, host_wide.cpp) for
// REGULAR rays (every 1/d finite).  Irregular rays keep the binary walk of the reference's own topology
// (trace.cuh), where NaNs from 0 * inf flow through the reference's ternaries box by box.
//
// Parity contract: for a regular ray the reference reports the closest primitive among those whose
// own leaf box AND primitive test pass (ties: lowest DFS slot) — inner boxes only decide how fast that set is
// found.  Here every child plane is decoded to the fp32 value the host verified,
//     dec(q) = fma(as_float(0x4B000000 | q), scale, base2),   dec(q_lo) <= exact lo,  dec(q_hi) >= exact hi,
// and tested with the reference's own fl(fl(plane - o) * inv).  That expression is monotone in `plane`, so a ray
// that passes an exact leaf box passes every decoded ancestor box: no candidate is lost.  The exact leaf box is
// tested (box_test_regular, bit-identical to BoundBox::IntersectRay for regular rays) whenever a primitive test
// accepts, so no candidate is gained either.  Pruning by distance uses the same slack rule as the binary walk.
//
// Walk (after Ylitie, Karras, Laine, HPG 2017): a stack entry is a node GROUP {first child index, hit bits of the
// inner children in octant order | imask}; the highest hit bit is the child to visit next.  Slot s of a node lies
// on the + side of axis a when bit a of s is set (host_wide.cpp), so for a ray whose direction is positive along
// the axes in `oct`, priority s ^ oct visits the children roughly front to back without sorting.
#pragma once
#include "trace.cuh"

namespace tutu {

struct WideRay {      // per-ray constants of the wide walk
  float2 nox, noy, noz;  // (-o, -o)
  float2 ix, iy, iz;     // (1/d, 1/d)
  unsigned oct;          // bit a set: d[a] >= 0
  bool neg_x, neg_y, neg_z;
};

__device__ __forceinline__ WideRay make_wide_ray(const RayPre& p, const Ray& r) {
  WideRay w;
  w.nox = make_float2(-p.ox, -p.ox), w.noy = make_float2(-p.oy, -p.oy), w.noz = make_float2(-p.oz, -p.oz);
  w.ix = make_float2(p.ix, p.ix), w.iy = make_float2(p.iy, p.iy), w.iz = make_float2(p.iz, p.iz);
  w.neg_x = r.dx < 0.f, w.neg_y = r.dy < 0.f, w.neg_z = r.dz < 0.f;
  w.oct = (w.neg_x ? 0u : 1u) | (w.neg_y ? 0u : 2u) | (w.neg_z ? 0u : 4u);
  return w;
}

// byte j of `word` -> as_float(0x4B000000 | byte) = 2^23 + byte (one PRMT)
template <int J>
__device__ __forceinline__ float wide_q(unsigned word) {
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540u | (unsigned)J));
}

// One child: (near, far) plane pair of each axis decoded with one FFMA2, then the reference's subtract and
// multiply as FADD2 / FMUL2 (add.rn(p, -o) is the IEEE p - o).  near = the plane the ray meets first on that
// axis (lo for d > 0), which is the min of the reference's two slab distances for a regular ray.
template <int J>
__device__ __forceinline__ bool wide_child_hit(const WideRay& w, unsigned nx, unsigned fx, unsigned ny, unsigned fy,
                                               unsigned nz, unsigned fz, float2 sx, float2 bx, float2 sy, float2 by,
                                               float2 sz, float2 bz, float lim) {
  const float2 px = __ffma2_rn(make_float2(wide_q<J>(nx), wide_q<J>(fx)), sx, bx);
  const float2 py = __ffma2_rn(make_float2(wide_q<J>(ny), wide_q<J>(fy)), sy, by);
  const float2 pz = __ffma2_rn(make_float2(wide_q<J>(nz), wide_q<J>(fz)), sz, bz);
  const float2 tx = __fmul2_rn(__fadd2_rn(px, w.nox), w.ix);
  const float2 ty = __fmul2_rn(__fadd2_rn(py, w.noy), w.iy);
  const float2 tz = __fmul2_rn(__fadd2_rn(pz, w.noz), w.iz);
  const float t_enter = fmaxf(tx.x, fmaxf(ty.x, tz.x));
  const float t_exit = fminf(tx.y, fminf(ty.y, tz.y));
  // t_enter <= t_exit && t_exit >= 0 (BoundBox.hpp:91) folded as in the small-scene test, plus the pruning limit
  return fmaxf(t_enter, 0.f) <= t_exit && t_enter <= lim;
}

// bit j of the result = the child in slot j is hit
__device__ __forceinline__ unsigned wide_node_hits(const WideRay& w, const uint4 h0, const uint4 h1, const uint4 q0,
                                                   const uint4 q1, const uint4 q2, const uint4 q3, float lim) {
  // h0 = {base2.x, base2.y, base2.z, scale.x}  h1 = {scale.y, scale.z, child_base, leaf_base}
  // q0 = {masks, qlo_x[0..3], qlo_x[4..7]}: see the byte layout in tutu_internal.hpp
  const float2 bx = make_float2(__uint_as_float(h0.x), __uint_as_float(h0.x));
  const float2 by = make_float2(__uint_as_float(h0.y), __uint_as_float(h0.y));
  const float2 bz = make_float2(__uint_as_float(h0.z), __uint_as_float(h0.z));
  const float2 sx = make_float2(__uint_as_float(h0.w), __uint_as_float(h0.w));
  const float2 sy = make_float2(__uint_as_float(h1.x), __uint_as_float(h1.x));
  const float2 sz = make_float2(__uint_as_float(h1.y), __uint_as_float(h1.y));
  // words: q0.z q0.w = qlo_x, q1.x q1.y = qlo_y, q1.z q1.w = qlo_z, q2.x q2.y = qhi_x, q2.z q2.w = qhi_y, q3.x q3.y = qhi_z
  const unsigned nx0 = w.neg_x ? q2.x : q0.z, nx1 = w.neg_x ? q2.y : q0.w, fx0 = w.neg_x ? q0.z : q2.x, fx1 = w.neg_x ? q0.w : q2.y;
  const unsigned ny0 = w.neg_y ? q2.z : q1.x, ny1 = w.neg_y ? q2.w : q1.y, fy0 = w.neg_y ? q1.x : q2.z, fy1 = w.neg_y ? q1.y : q2.w;
  const unsigned nz0 = w.neg_z ? q3.x : q1.z, nz1 = w.neg_z ? q3.y : q1.w, fz0 = w.neg_z ? q1.z : q3.x, fz1 = w.neg_z ? q1.w : q3.y;
  unsigned hits = 0u;
  if (wide_child_hit<0>(w, nx0, fx0, ny0, fy0, nz0, fz0, sx, bx, sy, by, sz, bz, lim)) hits |= 1u;
  if (wide_child_hit<1>(w, nx0, fx0, ny0, fy0, nz0, fz0, sx, bx, sy, by, sz, bz, lim)) hits |= 2u;
  if (wide_child_hit<2>(w, nx0, fx0, ny0, fy0, nz0, fz0, sx, bx, sy, by, sz, bz, lim)) hits |= 4u;
  if (wide_child_hit<3>(w, nx0, fx0, ny0, fy0, nz0, fz0, sx, bx, sy, by, sz, bz, lim)) hits |= 8u;
  if (wide_child_hit<0>(w, nx1, fx1, ny1, fy1, nz1, fz1, sx, bx, sy, by, sz, bz, lim)) hits |= 16u;
  if (wide_child_hit<1>(w, nx1, fx1, ny1, fy1, nz1, fz1, sx, bx, sy, by, sz, bz, lim)) hits |= 32u;
  if (wide_child_hit<2>(w, nx1, fx1, ny1, fy1, nz1, fz1, sx, bx, sy, by, sz, bz, lim)) hits |= 64u;
  if (wide_child_hit<3>(w, nx1, fx1, ny1, fy1, nz1, fz1, sx, bx, sy, by, sz, bz, lim)) hits |= 128u;
  return hits;
}

// bit s of x -> bit (s ^ oct): three conditional swaps of bit groups
__device__ __forceinline__ unsigned wide_permute(unsigned x, unsigned oct) {
  if (oct & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
  if (oct & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
  if (oct & 4u) x = ((x & 0x0Fu) << 4) | ((x >> 4) & 0x0Fu);
  return x;
}

__device__ __forceinline__ void wide_load(const uint4* __restrict__ n, uint4& h0, uint4& h1, uint4& q0, uint4& q1, uint4& q2,
                                          uint4& q3) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(h0.x), "=r"(h0.y), "=r"(h0.z), "=r"(h0.w), "=r"(h1.x), "=r"(h1.y), "=r"(h1.z), "=r"(h1.w)
               : "l"(n));
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w)
               : "l"(n + 2));
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w), "=r"(q3.x), "=r"(q3.y), "=r"(q3.z), "=r"(q3.w)
               : "l"(n + 4));
}

// Primitive test of wide leaf k (+ the exact leaf box when the primitive test accepts).  Returns true when an
// any-hit walk is over (blocked).
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool wide_leaf(const DevScene& sc, const Ray& r, const RayPre& p, float dis, unsigned k, Hit& best,
                                          VisitCount* vc) {
  const float4* g = sc.wleaf + 4 * (size_t)k;
  if (COUNT) vc->prims++;
  const uint32_t code = __float_as_uint(__ldg(g + 3).x);
  float t, u = 0.f, v = 0.f;
  bool hit;
  if (code & kSphereBit) {
    const SphereHit sh = sphere_test(g, r);
    hit = sh.hit;
    t = sh.t;
  } else {
    hit = tri_test(g, r, t, u, v);
  }
  if (!hit) return false;
  const int slot = (int)(code & kSlotMask);
  bool accept;
  if (ANY)
    accept = t < dis && !(fabsf(__fsub_rn(t, dis)) < 0.0001f);  // BVH.hpp:184
  else
    accept = t < best.t || (t == best.t && slot < (best.slot & (int)kSlotMask));  // BVH.hpp:165: lowest DFS leaf on ties
  if (!accept) return false;
  // the reference reaches this primitive only through its own leaf box (BVH.hpp:148,173)
  const float4 b0 = __ldg(sc.wbox + 2 * (size_t)k), b1 = __ldg(sc.wbox + 2 * (size_t)k + 1);
  float te;
  if (!box_test_regular(p, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, te)) return false;
  best.t = t;
  best.u = u;
  best.v = v;
  best.slot = (int)code;
  return ANY;
}

// Whole walk of one REGULAR ray whose root box is hit.  `st`: this thread's stack column (64-bit words; a ray
// pushes at most one group per level of the wide tree).  best must be initialised (miss record).
template <bool ANY, bool COUNT>
__device__ __forceinline__ void walk_wide(const DevScene& sc, const Ray& r, const RayPre& p, float dis, Hit& best,
                                          unsigned long long* stack, unsigned stride, VisitCount* vc) {
  const WideRay w = make_wide_ray(p, r);
  const uint4* __restrict__ nodes = reinterpret_cast<const uint4*>(sc.wide);
  unsigned gbase = 0u, gbits = 0x80000000u;  // node group: first child index, hit bits (31..24) | imask (7..0)
  int sp = 0;
  for (;;) {
    // ---- next inner child of the current group: highest priority bit ----
    const unsigned bit = 31u - (unsigned)__clz(gbits);
    gbits &= ~(1u << bit);
    const unsigned slot = (bit - 24u) ^ w.oct;
    const unsigned rel = (unsigned)__popc(gbits & 0xFFu & ((1u << slot) - 1u));
    const unsigned node = gbase + rel;
    if (gbits & 0xFF000000u) {
      stack[(unsigned)sp * stride] = ((unsigned long long)gbits << 32) | gbase;
      ++sp;
    }
    uint4 h0, h1, q0, q1, q2, q3;
    wide_load(nodes + 6 * (size_t)node, h0, h1, q0, q1, q2, q3);
    if (COUNT) vc->nodes++;
    const float limb = ANY ? dis : best.t;
    const float lim = fmaf(limb, sc.prune_rel, limb + sc.prune_abs);
    const unsigned hits = wide_node_hits(w, h0, h1, q0, q1, q2, q3, lim);
    const unsigned imask = q0.x & 0xFFu, lmask = (q0.x >> 8) & 0xFFu;
    gbase = h1.z;
    gbits = (wide_permute(hits & imask, w.oct) << 24) | imask;
    // ---- leaves of this node ----
    unsigned lhits = hits & lmask;
    while (lhits) {
      const unsigned s = (unsigned)__ffs((int)lhits) - 1u;
      lhits &= lhits - 1u;
      const unsigned k = h1.w + (unsigned)__popc(lmask & ((1u << s) - 1u));
      if (wide_leaf<ANY, COUNT>(sc, r, p, dis, k, best, vc)) return;
    }
    // ---- pop when this node contributed no inner child ----
    if (!(gbits & 0xFF000000u)) {
      if (sp == 0) return;
      --sp;
      const unsigned long long e = stack[(unsigned)sp * stride];
      gbase = (unsigned)e;
      gbits = (unsigned)(e >> 32);
    }
  }
}

// Rays the wide walk does not take (an infinite 1/d; scenes without a wide tree): the binary walk with a
// local-memory stack, kept out of line so that its stack arrays and registers do not weigh on the callers.
// TAG: one instantiation per calling kernel — cicc 12.9 segfaults when two kernels of a translation unit call
// the same instantiation of this function.
struct HitAny {
  Hit h;
  bool any;
};
template <bool ANY, int TAG>
static __device__ __noinline__ HitAny trace_binary_outofline(const DevScene& sc, const Ray r, const float dis) {
  HitAny out;
  out.any = traverse<ANY, 0, false>(sc, r, dis, out.h, nullptr);
  return out;
}

// Production entry of the queue / batch kernels on scenes with a tree: closest hit (ANY = false) or any hit
// within `dis`.  TAG = a number unique to the calling kernel (see trace_binary_outofline).  `stack` / `stride`: this thread's column of the block's shared-memory stack
// ((wide_depth + 1) 64-bit words per thread).  Returns "hit" / "blocked".
template <bool ANY, int TAG>
__device__ __forceinline__ bool trace_ray(const DevScene& sc, const Ray& r, float dis, Hit& best, unsigned long long* stack,
                                          unsigned stride) {
  best.t = FLT_MAX, best.u = 0.f, best.v = 0.f, best.slot = -1;
  if (sc.empty) return false;
  const RayPre p = make_pre(r);
  if (sc.wide != nullptr && ray_is_regular(p)) {
    float te;
    if (box_test_regular(p, sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2], te))
      walk_wide<ANY, false>(sc, r, p, dis, best, stack, stride, nullptr);
    return best.slot >= 0;
  }
  const HitAny o = trace_binary_outofline<ANY, TAG>(sc, r, dis);
  best = o.h;
  return o.any;
}

}  // namespace tutu
