// Host side of libtutu_b200: scene validation, the reference-compatible midpoint BVH build,
// flattening into the device layout, scene files and the synthetic ray-batch workload.
// Pure C++ (no CUDA) so the oracle harness can link it for scene-file IO.
//
// Built with -ffp-contract=off: everything that feeds hit parity (bounds, centroids, E1/E2/n,
// ray-generation constants) must round exactly like the reference's scalar float code.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>

#include "tutu_internal.hpp"

namespace tutu {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }

// ------------------------------------------------------------------------------------------
// reference float helpers (Vector.hpp:186,213-225)
// ------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 mul(float c, V3 v) { return {v.x * c, v.y * c, v.z * c}; }
static inline V3 divs(V3 v, float c) { return {v.x / c, v.y / c, v.z / c}; }
static inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline V3 normalized(V3 v) {
  float mag = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (mag > 0) {
    float inv = 1 / mag;
    return {v.x * inv, v.y * inv, v.z * inv};
  }
  return v;
}
static inline V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // Vector.hpp:186
static inline void st3(float* p, V3 v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}

// Object::initializeBound — Triangle.hpp:104-107 (BoundBox(v0,v1) then Union with v2, all
// fmin/fmax, BoundBox.hpp:13-27,112-124) and Sphere.hpp:129-133 (centre -/+ radius).
static Box prim_bound(const TutuPrim& p) {
  Box b;
  if (p.type == TUTU_PRIM_SPHERE) {
    float r = p.v[3];
    for (int a = 0; a < 3; ++a) {
      float lo = p.v[a] - r, hi = p.v[a] + r;
      b.lo[a] = fminf(lo, hi);
      b.hi[a] = fmaxf(lo, hi);
    }
    return b;
  }
  for (int a = 0; a < 3; ++a) {
    float lo = fminf(p.v[a], p.v[3 + a]);
    float hi = fmaxf(p.v[a], p.v[3 + a]);
    b.lo[a] = fminf(lo, p.v[6 + a]);
    b.hi[a] = fmaxf(hi, p.v[6 + a]);
  }
  return b;
}
static inline Box box_union(const Box& a, const Box& b) {  // BoundBox.hpp:97-109
  Box r;
  for (int i = 0; i < 3; ++i) {
    r.lo[i] = fminf(a.lo[i], b.lo[i]);
    r.hi[i] = fmaxf(a.hi[i], b.hi[i]);
  }
  return r;
}
static inline int max_extent(const Box& b) {  // BoundBox.hpp:43-52
  float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  if (dx > dy && dx > dz) return 0;
  if (dy > dz) return 1;
  return 2;
}

// chunks [begin, end) of [0, n) on up to 16 host threads (serial below 32 Ki items); fn(begin, end, thread index)
template <class F>
static void parallel_chunks(size_t n, F fn) {
  unsigned t = std::thread::hardware_concurrency();
  t = t < 1 ? 1 : (t > 16 ? 16 : t);
  if (n < (1u << 15) || t == 1) {
    fn((size_t)0, n, 0u);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(t);
  for (unsigned k = 0; k < t; ++k) pool.emplace_back([=] { fn(n * k / t, n * (k + 1) / t, k); });
  for (auto& th : pool) th.join();
}
constexpr unsigned kMaxHostThreads = 16;

// ------------------------------------------------------------------------------------------
// midpoint BVH, reference split rule (BVH.hpp:47-123)
// ------------------------------------------------------------------------------------------
namespace {
struct Builder {
  const std::vector<Box>& bounds;
  std::vector<V3> cent;  // BoundBox::Centroid(), BoundBox.hpp:35: 0.5*pMin + 0.5*pMax
  std::vector<uint32_t> idx;
  TutuBvhNode* nodes;

  Builder(const std::vector<Box>& b, TutuBvhNode* out) : bounds(b), nodes(out) {
    size_t n = b.size();
    cent.resize(n);
    idx.resize(n);
    for (size_t i = 0; i < n; ++i) {
      idx[i] = (uint32_t)i;
      cent[i] = {0.5f * b[i].lo[0] + 0.5f * b[i].hi[0], 0.5f * b[i].lo[1] + 0.5f * b[i].hi[1],
                 0.5f * b[i].lo[2] + 0.5f * b[i].hi[2]};
    }
  }

  // A subtree over n leaves always has 2n-1 nodes, so pre-order indices are known up front
  // (left = node+1, right = node+2*n_left) and big subtrees can be built on separate threads.
  void build(uint32_t node, uint32_t lo, uint32_t hi, int par_levels) {
    uint32_t n = hi - lo;
    if (n == 1) {
      nodes[node] = {-1, -1, (int32_t)idx[lo]};
      return;
    }
    if (n == 2) {  // BVH.hpp:61-67: no sort, list order kept
      nodes[node] = {(int32_t)node + 1, (int32_t)node + 2, -1};
      nodes[node + 1] = {-1, -1, (int32_t)idx[lo]};
      nodes[node + 2] = {-1, -1, (int32_t)idx[lo + 1]};
      return;
    }
    Box u = box_union(bounds[idx[lo]], bounds[idx[lo + 1]]);
    for (uint32_t i = lo + 2; i < hi; ++i) u = box_union(u, bounds[idx[i]]);
    const V3* c = cent.data();
    // Same std::sort, same comparator, same input sequence as BVH.hpp:85-103: the reference
    // sorts a copy holding exactly this subrange, so the (unstable) result is identical.
    switch (max_extent(u)) {
      case 0:
        std::sort(idx.begin() + lo, idx.begin() + hi,
                  [c](uint32_t a, uint32_t b) { return c[a].x < c[b].x; });
        break;
      case 1:
        std::sort(idx.begin() + lo, idx.begin() + hi,
                  [c](uint32_t a, uint32_t b) { return c[a].y < c[b].y; });
        break;
      default:
        std::sort(idx.begin() + lo, idx.begin() + hi,
                  [c](uint32_t a, uint32_t b) { return c[a].z < c[b].z; });
        break;
    }
    uint32_t nl = n / 2;  // BVH.hpp:108
    uint32_t mid = lo + nl;
    uint32_t left = node + 1, right = node + 2 * nl;
    nodes[node] = {(int32_t)left, (int32_t)right, -1};
    if (par_levels > 0 && n > (1u << 15)) {
      std::thread t([=] { build(left, lo, mid, par_levels - 1); });
      build(right, mid, hi, par_levels - 1);
      t.join();
    } else {
      build(left, lo, mid, 0);
      build(right, mid, hi, 0);
    }
  }
};
}  // namespace

static int build_bvh(const TutuPrim* prims, uint32_t n, std::vector<TutuBvhNode>* out) {
  out->clear();
  if (n == 0) return TUTU_OK;
  std::vector<Box> bounds(n);
  for (uint32_t i = 0; i < n; ++i) bounds[i] = prim_bound(prims[i]);
  out->resize(2 * (size_t)n - 1);
  Builder b(bounds, out->data());
  b.build(0, 0, n, 4);
  return TUTU_OK;
}

// ------------------------------------------------------------------------------------------
// ray-generation constants — Camera.hpp:12-48 + PathTracing.hpp:357-391
// ------------------------------------------------------------------------------------------
#define TUTU_M_PI 3.1415926535897f /* global.hpp:15 */
static inline float degree2Radians(float d) { return d * TUTU_M_PI / 180.f; }  // global.hpp:129

int compute_raygen(const TutuCamera* cam, RayGen* rg) {
  if (cam->width <= 0 || cam->height <= 0) {
    set_error("camera: width and height must be positive");
    return TUTU_E_INVALID;
  }
  // Camera::initialize
  V3 fwd = normalized(ld3(cam->viewdir));
  V3 right = normalized(cross(fwd, ld3(cam->updir)));
  V3 up = normalized(cross(right, fwd));
  float tanHalfHfov = tanf(degree2Radians(cam->hfov_deg * 0.5f));
  float imagePlaneDist = cam->width / (2.f * tanHalfHfov);
  // PathTracing::integrate
  V3 u = normalized(cross(fwd, up));
  V3 v = normalized(cross(u, fwd));
  float d = imagePlaneDist;
  if (cam->parallel_projection) d = 4.f;
  float width_half = fabsf(tanf(degree2Radians(cam->hfov_deg / 2.f)) * d);
  float aspect_ratio = cam->width / (float)cam->height;
  float height_half = width_half / aspect_ratio;
  V3 n = normalized(ld3(cam->viewdir));
  V3 eye = ld3(cam->eye);
  V3 base = add(eye, mul(d, n));
  V3 ul = add(sub(base, mul(width_half, u)), mul(height_half, v));
  V3 ur = add(add(base, mul(width_half, u)), mul(height_half, v));
  V3 ll = sub(sub(base, mul(width_half, u)), mul(height_half, v));
  V3 delta_h = {0, 0, 0}, delta_v = {0, 0, 0};
  if (cam->width != 1) delta_h = divs(sub(ur, ul), (float)(cam->width - 1));
  if (cam->height != 1) delta_v = divs(sub(ll, ul), (float)(cam->height - 1));
  V3 c_off_h = divs(sub(ur, ul), (float)(cam->width * 2));
  V3 c_off_v = divs(sub(ll, ul), (float)(cam->height * 2));
  st3(rg->eye, eye);
  st3(rg->ul, ul);
  st3(rg->delta_h, delta_h);
  st3(rg->delta_v, delta_v);
  st3(rg->c_off_h, c_off_h);
  st3(rg->c_off_v, c_off_v);
  rg->width = cam->width;
  rg->height = cam->height;
  return TUTU_OK;
}


// child boxes of a device inner node: per child and axis {lo, hi} pairs (trace.cuh: node_boxes)
static inline void set_node_boxes(InnerNode& in, const Box& l, const Box& r) {
  for (int a = 0; a < 3; ++a) {
    in.box[2 * a] = l.lo[a], in.box[2 * a + 1] = l.hi[a];
    in.box[6 + 2 * a] = r.lo[a], in.box[6 + 2 * a + 1] = r.hi[a];
  }
}

// ------------------------------------------------------------------------------------------
// traversal tree for regular rays (tutu_internal.hpp: FlatScene::inner_fast)
// ------------------------------------------------------------------------------------------
// For a ray whose three 1/d are finite the reference's answer depends only on the LEAF boxes and
// the primitives (slab test monotone in the box + every reference inner box is the exact fmin/fmax
// union of its leaves: "leaf box hit => every ancestor box hit").  Any binary tree
// over the same leaves whose inner boxes are exact unions of their leaves' boxes therefore returns
// bit-identical hits, and the device is free to walk a better one than the reference's
// median-by-count split: a binned surface-area-heuristic tree (16 bins, 3 axes), depth limited so
// that the traversal stacks stay valid.  Leaf refs keep the reference's DFS slot numbers, which is
// what the equal-t tie rule compares.  Irregular rays keep walking the reference topology.
namespace {
struct FastBuilder {
  // The builder partitions 28-byte records {leaf box, leaf id} in place instead of an index array: every pass
  // over a node's primitives is then a sequential scan (with indices the 10^6-leaf build spent its time on cache
  // misses of box[idx[k]]).  The centroid is recomputed from the box.
  struct Rec {
    Box box;
    uint32_t id;
  };
  const PodVec<uint32_t>& code;
  PodVec<Rec> rec;
  InnerNode* out;
  int max_depth;
  uint32_t depth_seen = 0;

  static float area(const Box& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.f * (dx * dy + dy * dz + dz * dx);
  }
  static Box empty_box() {
    Box b;
    for (int a = 0; a < 3; ++a) b.lo[a] = FLT_MAX, b.hi[a] = -FLT_MAX;
    return b;
  }
  static void grow(Box& b, const Box& o) {
    for (int a = 0; a < 3; ++a) {
      b.lo[a] = fminf(b.lo[a], o.lo[a]);  // fmin/fmax like BoundBox::Union (BoundBox.hpp:97-124): exact
      b.hi[a] = fmaxf(b.hi[a], o.hi[a]);
    }
  }
  static V3 centroid(const Box& b) { return {0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2])}; }
  static int ceil_log2(uint32_t n) {
    int l = 0;
    while ((1ull << l) < n) ++l;
    return l;
  }

  // builds the subtree over idx[first, first+count); its count-1 inner nodes occupy out[base, base+count-1)
  // in pre-order.  Returns the child ref and the exact union box.
  int32_t build(uint32_t first, uint32_t count, uint32_t base, int depth, Box* box_out, int par_levels, uint32_t* deepest) {
    if (count == 1) {
      *box_out = rec[first].box;
      return (int32_t)~code[rec[first].id];
    }
    if ((uint32_t)depth + 1 > *deepest) *deepest = (uint32_t)depth + 1;
    uint32_t nl = count / 2;
    bool median = (max_depth - depth) <= ceil_log2(count) + 1;
    if (!median) {
      // Centroid bounds and the 3 x 16 bins in one pass each; nodes of more than 128 Ki primitives split both
      // passes over the host threads (min / max and counts merge exactly, so the tree does not depend on it).
      const unsigned workers = (par_levels > 0 && count > (1u << 17)) ? std::min(16u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
      auto chunk = [&](unsigned w) { return std::pair<uint32_t, uint32_t>(first + (uint32_t)((uint64_t)count * w / workers), first + (uint32_t)((uint64_t)count * (w + 1) / workers)); };
      auto for_workers = [&](auto fn) {
        if (workers == 1) {
          fn(0u);
          return;
        }
        std::vector<std::thread> pool;
        for (unsigned w = 0; w < workers; ++w) pool.emplace_back([&fn, w] { fn(w); });
        for (auto& t : pool) t.join();
      };
      Box cb_one = empty_box();  // no heap allocation on the (million-fold) single-worker path
      std::vector<Box> cb_many;
      if (workers > 1) cb_many.assign(workers, empty_box());
      Box* cbs = workers > 1 ? cb_many.data() : &cb_one;
      for_workers([&](unsigned w) {
        Box c0 = empty_box();
        const auto r = chunk(w);
        for (uint32_t k = r.first; k < r.second; ++k) {
          const V3 c = centroid(rec[k].box);
          c0.lo[0] = fminf(c0.lo[0], c.x), c0.hi[0] = fmaxf(c0.hi[0], c.x);
          c0.lo[1] = fminf(c0.lo[1], c.y), c0.hi[1] = fmaxf(c0.hi[1], c.y);
          c0.lo[2] = fminf(c0.lo[2], c.z), c0.hi[2] = fmaxf(c0.hi[2], c.z);
        }
        cbs[w] = c0;
      });
      Box cb = cbs[0];
      for (unsigned w = 1; w < workers; ++w) grow(cb, cbs[w]);
      constexpr int NB = 16;
      struct Bins {
        uint32_t cnt[3][NB];
        Box bb[3][NB];
      };
      float scale3[3];
      bool use_axis[3];
      for (int a = 0; a < 3; ++a) {
        const float ext = cb.hi[a] - cb.lo[a];
        use_axis[a] = ext > 0.f;
        scale3[a] = use_axis[a] ? NB / ext : 0.f;
      }
      Bins bins_one;
      std::vector<Bins> bins_many;
      if (workers > 1) bins_many.resize(workers);
      Bins* bins = workers > 1 ? bins_many.data() : &bins_one;
      for_workers([&](unsigned w) {
        Bins& B = bins[w];
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < NB; ++k) B.cnt[a][k] = 0, B.bb[a][k] = empty_box();
        const auto r = chunk(w);
        for (uint32_t k = r.first; k < r.second; ++k) {
          const Box& lb = rec[k].box;
          const V3 cc = centroid(lb);
          const float c3[3] = {cc.x, cc.y, cc.z};
          for (int a = 0; a < 3; ++a) {
            if (!use_axis[a]) continue;
            int bk = (int)((c3[a] - cb.lo[a]) * scale3[a]);
            bk = bk < 0 ? 0 : (bk >= NB ? NB - 1 : bk);
            B.cnt[a][bk]++;
            grow(B.bb[a][bk], lb);
          }
        }
      });
      for (unsigned w = 1; w < workers; ++w)
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < NB; ++k) {
            bins[0].cnt[a][k] += bins[w].cnt[a][k];
            if (bins[w].cnt[a][k]) grow(bins[0].bb[a][k], bins[w].bb[a][k]);
          }
      float best = FLT_MAX;
      int best_axis = -1, best_split = 0;
      for (int a = 0; a < 3; ++a) {
        if (!use_axis[a]) continue;
        const uint32_t* cnt = bins[0].cnt[a];
        const Box* bb = bins[0].bb[a];
        float right_area[NB];
        uint32_t right_cnt[NB];
        Box acc = empty_box();
        uint32_t n = 0;
        for (int b = NB - 1; b > 0; --b) {
          if (cnt[b]) grow(acc, bb[b]);
          n += cnt[b];
          right_area[b] = n ? area(acc) : 0.f;
          right_cnt[b] = n;
        }
        acc = empty_box();
        n = 0;
        for (int b = 0; b < NB - 1; ++b) {
          if (cnt[b]) grow(acc, bb[b]);
          n += cnt[b];
          if (n == 0 || right_cnt[b + 1] == 0) continue;
          const float cost = area(acc) * n + right_area[b + 1] * right_cnt[b + 1];
          if (cost < best) best = cost, best_axis = a, best_split = b + 1;
        }
      }
      if (best_axis >= 0) {
        const int a = best_axis;
        const float scale = NB / (cb.hi[a] - cb.lo[a]);
        auto mid = std::partition(rec.begin() + first, rec.begin() + first + count, [&](const Rec& r) {
          const V3 cc = centroid(r.box);
          const float c = a == 0 ? cc.x : a == 1 ? cc.y : cc.z;
          int b = (int)((c - cb.lo[a]) * scale);
          b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
          return b < best_split;
        });
        nl = (uint32_t)(mid - (rec.begin() + first));
        if (nl == 0 || nl == count) nl = count / 2;  // cannot happen; keeps the recursion well founded
      }
    }
    Box bl, br;
    int32_t lref, rref;
    uint32_t dl = 0, dr = 0;
    if (par_levels > 0 && count > (1u << 15)) {
      std::thread t([&] { lref = build(first, nl, base + 1, depth + 1, &bl, par_levels - 1, &dl); });
      rref = build(first + nl, count - nl, base + nl, depth + 1, &br, par_levels - 1, &dr);
      t.join();
    } else {
      lref = build(first, nl, base + 1, depth + 1, &bl, 0, &dl);
      rref = build(first + nl, count - nl, base + nl, depth + 1, &br, 0, &dr);
    }
    if (dl > *deepest) *deepest = dl;
    if (dr > *deepest) *deepest = dr;
    InnerNode& in = out[base];
    set_node_boxes(in, bl, br);
    in.left = lref;
    in.right = rref;
    in.pad0 = in.pad1 = 0;
    *box_out = bl;
    grow(*box_out, br);
    return (int32_t)base;
  }
};
}  // namespace

// leaf_code[s] = slot | SPHERE_BIT.  On return fs->inner_fast / root_ref_fast / depth_fast are set;
// when the leaf boxes are not all finite (or the tree could not respect the depth limit) the
// reference topology is reused.
static void build_fast_tree(FlatScene* fs, const PodVec<uint32_t>& leaf_code, int max_depth) {
  const uint32_t n = (uint32_t)fs->leaf_box.size();
  fs->inner_fast = fs->inner;
  fs->root_ref_fast = fs->root_ref;
  fs->depth_fast = fs->depth;
  if (n < 3) return;
#ifdef TUTU_EXPERIMENTS
  if (getenv("TUTU_NO_FAST_TREE")) return;
#endif
  for (const Box& b : fs->leaf_box)
    for (int a = 0; a < 3; ++a)
      if (!std::isfinite(b.lo[a]) || !std::isfinite(b.hi[a])) return;
  if (FastBuilder::ceil_log2(n) + 2 > max_depth) return;
  PodVec<InnerNode> out(n - 1);
  FastBuilder fb{leaf_code, {}, out.data(), max_depth};
  fb.rec.resize(n);
  for (uint32_t i = 0; i < n; ++i) fb.rec[i] = {fs->leaf_box[i], i};
  Box root;
  uint32_t deepest = 0;
  const int32_t ref = fb.build(0, n, 0, 0, &root, 4, &deepest);
  if ((int)deepest > max_depth) return;
  fs->inner_fast = std::move(out);
  fs->root_ref_fast = ref;
  fs->depth_fast = deepest;
}

void build_host_fast_tree(FlatScene* fs) { build_fast_tree(fs, fs->leaf_code, kFastTreeMaxDepth); }

// ------------------------------------------------------------------------------------------
// BDPT camera constants — Camera.hpp:12-48, Vector.hpp:228-372, BDPT.hpp:396-418
// ------------------------------------------------------------------------------------------
namespace {
struct M4 {
  float e[16];
  M4() {
    for (float& v : e) v = 0.f;
  }
  float get(int r, int c) const { return e[c + r * 4]; }
  void set(int r, int c, float v) { e[c + r * 4] = v; }
  void row(int r, float a, float b, float c, float d) { e[r * 4] = a, e[r * 4 + 1] = b, e[r * 4 + 2] = c, e[r * 4 + 3] = d; }
};
M4 mmul(const M4& l, const M4& r) {  // Vector.hpp:337-349 (accumulates from 0, left to right)
  M4 res;
  for (int row = 0; row < 4; row++)
    for (int col = 0; col < 4; col++) {
      float acc = 0;
      for (int i = 0; i < 4; i++) acc += l.get(row, i) * r.get(i, col);
      res.set(row, col, acc);
    }
  return res;
}
}  // namespace

int compute_bdpt_cam(const TutuCamera* cam, BdptCamConsts* out) {
  if (cam->width <= 0 || cam->height <= 0) {
    set_error("camera: width and height must be positive");
    return TUTU_E_INVALID;
  }
  // Camera::initialize
  V3 position = ld3(cam->eye);
  V3 fwd = normalized(ld3(cam->viewdir));
  V3 right = normalized(cross(fwd, ld3(cam->updir)));
  V3 up = normalized(cross(right, fwd));
  V3 nfwd = {-fwd.x, -fwd.y, -fwd.z};
  V3 pos = {dot(right, position), dot(up, position), dot(nfwd, position)};
  M4 world2Cam;
  world2Cam.row(0, right.x, right.y, right.z, -pos.x);
  world2Cam.row(1, up.x, up.y, up.z, -pos.y);
  world2Cam.row(2, nfwd.x, nfwd.y, nfwd.z, -pos.z);
  world2Cam.row(3, 0.f, 0.f, 0.f, 1.f);
  const float aNear = 0.1f, aFar = 10000.f, aspect = (float)cam->width / cam->height;
  M4 p2o;  // getPerspectiveMatrix, Vector.hpp:352-372
  p2o.e[0] = aNear, p2o.e[5] = aNear, p2o.e[10] = (aNear + aFar), p2o.e[11] = aNear * aFar, p2o.e[14] = -1.0f;
  float r = tanf(((float)cam->hfov_deg / 2) * 3.1415926535897f / 180) * aNear;
  float l = -r;
  float t = r / aspect;
  float b = -t;
  M4 orth_trans, orth_scale;
  orth_trans.row(0, 1, 0, 0, -(r + l) / 2);
  orth_trans.row(1, 0, 1, 0, -(t + b) / 2);
  orth_trans.row(2, 0, 0, 1, -(aNear + aFar) / 2);
  orth_trans.row(3, 0, 0, 0, 1);
  orth_scale.row(0, 2 / (r - l), 0, 0, 0);
  orth_scale.row(1, 0, 2 / -(t - b), 0, 0);
  orth_scale.row(2, 0, 0, 2 / (aNear - aFar), 0);
  orth_scale.row(3, 0, 0, 0, 1);
  M4 perspective = mmul(mmul(orth_scale, orth_trans), p2o);
  M4 world2ndc = mmul(perspective, world2Cam);
  M4 translate;  // Mat4f::getTranslate(1,1,0)
  translate.set(0, 3, 1.f), translate.set(1, 3, 1.f), translate.set(2, 3, 0.f);
  translate.set(0, 0, 1), translate.set(1, 1, 1), translate.set(2, 2, 1), translate.set(3, 3, 1);
  M4 scale;  // Mat4f::getScale(w/2, h/2, 0)
  scale.set(3, 3, 1), scale.set(0, 0, cam->width * 0.5f), scale.set(1, 1, cam->height * 0.5f), scale.set(2, 2, 0);
  M4 world2Raster = mmul(scale, mmul(translate, world2ndc));
  memcpy(out->w2r, world2Raster.e, sizeof(out->w2r));
  float tanHalfHfov = tanf(degree2Radians(cam->hfov_deg * 0.5f));
  out->imagePlaneDist = cam->width / (2.f * tanHalfHfov);
  out->filmPlaneAreaInv = 1.f / (cam->width * cam->height);
  out->lensAreaInv = 1.f;
  // BDPT::integrate (no parallel-projection override there)
  V3 u = normalized(cross(fwd, up));
  V3 v = normalized(cross(u, fwd));
  float d = out->imagePlaneDist;
  float width_half = fabsf(tanf(degree2Radians(cam->hfov_deg / 2.f)) * d);
  float aspect_ratio = cam->width / (float)cam->height;
  float height_half = width_half / aspect_ratio;
  V3 n = normalized(ld3(cam->viewdir));
  V3 base = add(position, mul(d, n));
  V3 ul = add(sub(base, mul(width_half, u)), mul(height_half, v));
  V3 ur = add(add(base, mul(width_half, u)), mul(height_half, v));
  V3 ll = sub(sub(base, mul(width_half, u)), mul(height_half, v));
  V3 delta_h = {0, 0, 0}, delta_v = {0, 0, 0};
  if (cam->width != 1) delta_h = divs(sub(ur, ul), (float)(cam->width - 1));
  if (cam->height != 1) delta_v = divs(sub(ll, ul), (float)(cam->height - 1));
  V3 c_off_h = divs(sub(ur, ul), (float)(cam->width * 2));
  V3 c_off_v = divs(sub(ll, ul), (float)(cam->height * 2));
  st3(out->eye, position);
  st3(out->fwd, fwd);
  st3(out->ul, ul);
  st3(out->dh, delta_h);
  st3(out->dv, delta_v);
  st3(out->coh, c_off_h);
  st3(out->cov, c_off_v);
  out->width = cam->width;
  out->height = cam->height;
  return TUTU_OK;
}

// ------------------------------------------------------------------------------------------
// Postprocessor::getGaussianBlurTexture's weights (Postprocessor.hpp:75-79): the lambda
//   (1 / sqrt(2 * M_PI * standardDev)) * pow(E, -(inputX * inputX) / (2 * standardDev * standardDev))
// with M_PI = 3.1415926535897f (global.hpp:15) and static float E = 2.7182818f: every operand is a float, so
// the C++ overloads are the float ones (sqrtf, powf).  kernelSum adds the taps in tap order.
// ------------------------------------------------------------------------------------------
void post_gaussian_weights(int kernel_size, float stddev, float* g, float* sum, int* start) {
  const float E = 2.7182818f;
  const int s0 = (int)(-kernel_size * 0.5);  // int startY = -kernelSize * 0.5;
  float acc = 0;
  for (int i = 0; i < kernel_size; ++i) {
    const int x = s0 + i;
    g[i] = (1 / sqrtf(2 * TUTU_M_PI * stddev)) * powf(E, -(x * x) / (2 * stddev * stddev));
    acc += g[i];
  }
  *sum = acc;
  *start = s0;
}

// ------------------------------------------------------------------------------------------
// flatten
// ------------------------------------------------------------------------------------------
static inline bool has_emission(const TutuMaterial& m) {  // Material.hpp:54-56
  return m.emission[0] || m.emission[1] || m.emission[2];
}

int flatten_scene(const TutuSceneDesc* desc, FlatScene* fs, bool host_fast_tree) {
  if (!desc || desc->struct_size != sizeof(TutuSceneDesc)) {
    set_error("scene: struct_size mismatch (ABI version skew?)");
    return TUTU_E_INVALID;
  }
  const uint32_t n = desc->n_prims;
  if (n >= SPHERE_BIT) {
    set_error("scene: too many primitives");
    return TUTU_E_INVALID;
  }
  if ((n && !desc->prims) || (desc->n_materials && !desc->materials)) {
    set_error("scene: null prims/materials");
    return TUTU_E_INVALID;
  }
  for (int c = 0; c < 4; ++c) {
    if (desc->n_tex[c] && !desc->tex[c]) {
      set_error("scene: null texture list");
      return TUTU_E_INVALID;
    }
    for (uint32_t i = 0; i < desc->n_tex[c]; ++i) {
      const TutuTexture& t = desc->tex[c][i];
      if (t.width < 0 || t.height < 0 || ((size_t)t.width * t.height != 0 && !t.rgb)) {
        set_error("scene: malformed texture");
        return TUTU_E_INVALID;
      }
    }
  }
  for (uint32_t i = 0; i < n; ++i) {
    const TutuPrim& p = desc->prims[i];
    if (p.type != TUTU_PRIM_TRIANGLE && p.type != TUTU_PRIM_SPHERE) {
      set_error("scene: unknown primitive type at index " + std::to_string(i));
      return TUTU_E_INVALID;
    }
    if (p.material < 0 || (uint32_t)p.material >= desc->n_materials) {
      set_error("scene: material index out of range at prim " + std::to_string(i));
      return TUTU_E_INVALID;
    }
    if (p.tex_active) {
      // IIntegrator.hpp:92-96,108-112,119-123: an out-of-range map index is a hard exit(1)
      // in the reference; here it is an upload error.
      const int32_t ids[4] = {p.tex_diffuse, p.tex_normal, p.tex_roughness, p.tex_metallic};
      for (int c = 0; c < 4; ++c)
        if (ids[c] < -1 || (ids[c] >= 0 && (uint32_t)ids[c] >= desc->n_tex[c])) {
          set_error("scene: texture index out of range at prim " + std::to_string(i));
          return TUTU_E_INVALID;
        }
    }
  }

  *fs = FlatScene();
  fs->n_prims = n;
  fs->camera = desc->camera;
  memcpy(fs->bkgcolor, desc->bkgcolor, sizeof(fs->bkgcolor));
  fs->eta = desc->eta;
  int rc = compute_raygen(&desc->camera, &fs->raygen);
  if (rc != TUTU_OK) return rc;
  rc = compute_bdpt_cam(&desc->camera, &fs->bdpt_cam);
  if (rc != TUTU_OK) return rc;

  // materials
  fs->materials.resize(desc->n_materials);
  for (uint32_t i = 0; i < desc->n_materials; ++i) {
    const TutuMaterial& m = desc->materials[i];
    DevMaterial& d = fs->materials[i];
    memset(&d, 0, sizeof(d));
    memcpy(d.diffuse, m.diffuse, 12);
    memcpy(d.specular, m.specular, 12);
    memcpy(d.emission, m.emission, 12);
    d.type = m.type;
    d.alpha = m.alpha;
    d.eta = m.eta;
    d.roughness = m.roughness;
    d.metallic = m.metallic;
    d.has_emission = has_emission(m) ? 1 : 0;
  }

  // textures: pooled float4 texels
  for (int c = 0; c < 4; ++c) {
    fs->tex_headers[c].resize(desc->n_tex[c]);
    for (uint32_t i = 0; i < desc->n_tex[c]; ++i) {
      const TutuTexture& t = desc->tex[c][i];
      TexHeader h;
      h.width = t.width;
      h.height = t.height;
      h.offset = (uint32_t)(fs->texels.size() / 4);
      h.n_texels = (uint32_t)((size_t)t.width * t.height);
      fs->tex_headers[c][i] = h;
      size_t base = fs->texels.size();
      fs->texels.resize(base + (size_t)h.n_texels * 4);
      for (size_t k = 0; k < h.n_texels; ++k) {
        fs->texels[base + 4 * k + 0] = t.rgb[3 * k + 0];
        fs->texels[base + 4 * k + 1] = t.rgb[3 * k + 1];
        fs->texels[base + 4 * k + 2] = t.rgb[3 * k + 2];
        fs->texels[base + 4 * k + 3] = 0.f;
      }
    }
  }

  // lights, objList order (PPMGenerator.hpp:317-324)
  if (n == 0) {
    fs->empty = true;
    return TUTU_OK;
  }
  fs->empty = false;

  // tree: given (reference export) or built here
  std::vector<TutuBvhNode> built;
  const TutuBvhNode* nodes = desc->bvh_nodes;
  uint32_t n_nodes = desc->n_bvh_nodes;
  if (!nodes || n_nodes == 0) {
    build_bvh(desc->prims, n, &built);
    nodes = built.data();
    n_nodes = (uint32_t)built.size();
  }
  if (n_nodes != 2 * n - 1) {
    set_error("scene: bvh node count must be 2*n_prims-1");
    return TUTU_E_INVALID;
  }
  fs->n_ref_nodes = n_nodes;

  // explicit DFS (left first) -> visit order, leaf slots, inner indices
  std::vector<uint32_t> order;
  order.reserve(n_nodes);
  std::vector<int32_t> inner_of(n_nodes, -1), slot_of(n_nodes, -1);
  std::vector<uint32_t> depth_of(n_nodes, 0);
  std::vector<uint8_t> seen(n_nodes, 0), prim_seen(n, 0);
  {
    std::vector<uint32_t> stack;
    stack.push_back(0);
    uint32_t n_inner = 0, n_slot = 0;
    while (!stack.empty()) {
      uint32_t i = stack.back();
      stack.pop_back();
      if (i >= n_nodes || seen[i]) {
        set_error("scene: bvh is not a tree");
        return TUTU_E_INVALID;
      }
      seen[i] = 1;
      order.push_back(i);
      const TutuBvhNode& nd = nodes[i];
      bool leaf = nd.left < 0 && nd.right < 0;
      if (leaf) {
        if (nd.prim < 0 || (uint32_t)nd.prim >= n || prim_seen[nd.prim]) {
          set_error("scene: bvh leaf has a bad or duplicated prim index");
          return TUTU_E_INVALID;
        }
        prim_seen[nd.prim] = 1;
        slot_of[i] = (int32_t)n_slot++;
      } else {
        if (nd.left < 0 || nd.right < 0 || (uint32_t)nd.left >= n_nodes ||
            (uint32_t)nd.right >= n_nodes) {
          set_error("scene: bvh inner node needs two children");
          return TUTU_E_INVALID;
        }
        inner_of[i] = (int32_t)n_inner++;
        depth_of[nd.left] = depth_of[nd.right] = depth_of[i] + 1;
        fs->depth = std::max(fs->depth, depth_of[i] + 1);
        stack.push_back((uint32_t)nd.right);
        stack.push_back((uint32_t)nd.left);  // left popped first
      }
    }
    if (n_slot != n || n_inner != n - 1) {
      set_error("scene: bvh does not cover every primitive exactly once");
      return TUTU_E_INVALID;
    }
  }

  // node boxes bottom-up: children always come later than their parent in `order`
  PodVec<Box> nbox(n_nodes);
  parallel_chunks(n_nodes, [&](size_t b, size_t e, unsigned) {  // leaf boxes are independent of each other
    for (size_t i = b; i < e; ++i)
      if (nodes[i].left < 0) nbox[i] = prim_bound(desc->prims[nodes[i].prim]);
  });
  {
    // inner boxes level by level from the deepest one (the nodes of a level are independent of each other)
    std::vector<uint32_t> level_start(fs->depth + 2, 0);
    for (uint32_t i = 0; i < n_nodes; ++i)
      if (nodes[i].left >= 0) level_start[depth_of[i] + 1]++;
    for (size_t d = 1; d < level_start.size(); ++d) level_start[d] += level_start[d - 1];
    PodVec<uint32_t> by_level(n - 1);
    {
      std::vector<uint32_t> cursor(level_start.begin(), level_start.end() - 1);
      for (uint32_t i = 0; i < n_nodes; ++i)
        if (nodes[i].left >= 0) by_level[cursor[depth_of[i]]++] = i;
    }
    for (size_t d = level_start.size() - 1; d-- > 0;) {
      const uint32_t first = level_start[d], count = level_start[d + 1] - first;
      parallel_chunks(count, [&](size_t b, size_t e, unsigned) {
        for (size_t k = b; k < e; ++k) {
          const uint32_t i = by_level[first + k];
          nbox[i] = box_union(nbox[nodes[i].left], nbox[nodes[i].right]);  // BVH.hpp:65,119
        }
      });
    }
  }
  fs->root_box = nbox[0];

  auto ref_of = [&](int32_t child) -> int32_t {
    if (nodes[child].left >= 0) return inner_of[child];
    uint32_t code = (uint32_t)slot_of[child];
    if (desc->prims[nodes[child].prim].type == TUTU_PRIM_SPHERE) code |= SPHERE_BIT;
    return (int32_t)~code;
  };
  fs->root_ref = ref_of(0);

  fs->inner.resize(n - 1);
  fs->geom.resize(n);
  fs->shade.resize(n);
  fs->slot_to_prim.resize(n);
  fs->leaf_box.resize(n);
  bool any_tex = false;
  for (uint32_t i = 0; i < n; ++i) any_tex |= desc->prims[i].tex_active != 0;
  if (any_tex) fs->leaftex.resize(n);

  float max_edge_of[kMaxHostThreads] = {};
  parallel_chunks(n_nodes, [&](size_t chunk_begin, size_t chunk_end, unsigned tid) {
  float max_edge = 0.f;  // per thread, combined below
  for (size_t i = chunk_begin; i < chunk_end; ++i) {
    const TutuBvhNode& nd = nodes[i];
    if (nd.left >= 0) {
      InnerNode& in = fs->inner[inner_of[i]];
      set_node_boxes(in, nbox[nd.left], nbox[nd.right]);
      in.left = ref_of(nd.left);
      in.right = ref_of(nd.right);
      in.pad0 = in.pad1 = 0;
      continue;
    }
    const uint32_t s = (uint32_t)slot_of[i];
    const TutuPrim& p = desc->prims[nd.prim];
    fs->slot_to_prim[s] = nd.prim;
    fs->leaf_box[s] = nbox[i];
    LeafGeom& g = fs->geom[s];
    LeafShade& sh = fs->shade[s];
    memset(&g, 0, sizeof(g));
    memset(&sh, 0, sizeof(sh));
    sh.flags = (uint32_t)p.material & MAT_MASK;
    if (p.tex_active) sh.flags |= TEX_ACTIVE_BIT;
    if (p.type == TUTU_PRIM_SPHERE) {
      g.f[0] = p.v[0];
      g.f[1] = p.v[1];
      g.f[2] = p.v[2];
      g.f[3] = p.v[3];
      if (2.f * fabsf(p.v[3]) > max_edge && fabsf(p.v[3]) < 1.0e38f) max_edge = 2.f * fabsf(p.v[3]);
      sh.flags |= SHADE_SPHERE_BIT;
    } else {
      // Triangle.hpp:25-35: E1, E2 and the unit geometric normal are ray independent, so they
      // are evaluated once here with the very same float operations.
      V3 v0 = ld3(p.v), v1 = ld3(p.v + 3), v2 = ld3(p.v + 6);
      V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
      V3 nn = normalized(cross(e1, e2));
      {
        V3 e3 = sub(v2, v1);
        float l1 = sqrtf(e1.x * e1.x + e1.y * e1.y + e1.z * e1.z);
        float l2 = sqrtf(e2.x * e2.x + e2.y * e2.y + e2.z * e2.z);
        float l3 = sqrtf(e3.x * e3.x + e3.y * e3.y + e3.z * e3.z);
        float l = fmaxf(l1, fmaxf(l2, l3));
        if (l > max_edge && l < 3.0e38f) max_edge = l;
      }
      st3(g.f + 0, v0);
      st3(g.f + 3, e1);
      st3(g.f + 6, e2);
      st3(g.f + 9, nn);
      sh.f[0] = p.n[0], sh.f[1] = p.n[1], sh.f[2] = p.n[2], sh.f[3] = p.uv[0];
      sh.f[4] = p.n[3], sh.f[5] = p.n[4], sh.f[6] = p.n[5], sh.f[7] = p.uv[1];
      sh.f[8] = p.n[6], sh.f[9] = p.n[7], sh.f[10] = p.n[8], sh.f[11] = p.uv[2];
      sh.f[12] = p.uv[3], sh.f[13] = p.uv[4], sh.f[14] = p.uv[5];
    }
    if (any_tex) {
      LeafTex& lt = fs->leaftex[s];
      lt.diffuse = p.tex_active ? p.tex_diffuse : -1;
      lt.normal = p.tex_active ? p.tex_normal : -1;
      lt.roughness = p.tex_active ? p.tex_roughness : -1;
      lt.metallic = p.tex_active ? p.tex_metallic : -1;
    }
  }
  max_edge_of[tid] = max_edge;
  });
  for (float m : max_edge_of) fs->max_edge = fmaxf(fs->max_edge, m);

  // slot lookup for lights
  bool any_light = false;
  for (uint32_t k = 0; k < desc->n_materials; ++k) any_light = any_light || has_emission(desc->materials[k]);
  std::vector<int32_t> prim_slot;
  if (any_light) {
    prim_slot.assign(n, -1);
    for (uint32_t s = 0; s < n; ++s) prim_slot[fs->slot_to_prim[s]] = (int32_t)s;
  }
  for (uint32_t i = 0; any_light && i < n; ++i) {
    const TutuPrim& p = desc->prims[i];
    const TutuMaterial& m = desc->materials[p.material];
    if (!has_emission(m)) continue;
    DevLight L;
    memset(&L, 0, sizeof(L));
    L.type = p.type;
    L.slot = prim_slot[i];
    L.material = p.material;
    memcpy(L.emission, m.emission, 12);
    if (p.type == TUTU_PRIM_SPHERE) {
      memcpy(L.v0, p.v, 12);
      L.radius = p.v[3];
      L.area = p.v[3] * p.v[3] * TUTU_M_PI;  // Sphere.hpp:135-137 (sic: pi r^2)
    } else {
      memcpy(L.v0, p.v, 12);
      memcpy(L.v1, p.v + 3, 12);
      memcpy(L.v2, p.v + 6, 12);
      memcpy(L.n0, p.n, 12);
      memcpy(L.n1, p.n + 3, 12);
      memcpy(L.n2, p.n + 6, 12);
      V3 c = cross(sub(ld3(p.v + 3), ld3(p.v)), sub(ld3(p.v + 6), ld3(p.v)));
      L.area = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z) * 0.5f;  // Triangle.hpp:109-116
    }
    fs->lights.push_back(L);
  }

  // traversal tree for regular rays (the reference topology stays in fs->inner for the others)
  fs->leaf_code.resize(n);
  for (uint32_t s = 0; s < n; ++s) fs->leaf_code[s] = s | ((fs->shade[s].flags & SHADE_SPHERE_BIT) ? SPHERE_BIT : 0u);
  if (host_fast_tree) {
    build_fast_tree(fs, fs->leaf_code, kFastTreeMaxDepth);
  } else {  // the caller replaces this with a device-built tree (or calls build_host_fast_tree on failure)
    fs->inner_fast.clear();
    fs->root_ref_fast = fs->root_ref;
    fs->depth_fast = fs->depth;
  }
  return TUTU_OK;
}

}  // namespace tutu

// ==========================================================================================
// extern "C": host-only entry points
// ==========================================================================================
using namespace tutu;

extern "C" int tutu_abi_version(void) { return TUTU_ABI_VERSION; }

extern "C" void tutu_post_params_default(TutuPostParams* out) {  // Postprocessor.hpp:10-14, :141
  if (!out) return;
  out->emissive_norm = 3.f;
  out->strength = 2.f;
  out->gaussian_loops = 1;
  out->kernel_size = 10;
  out->stddev = 30.f;
  out->exposure = 1.5f;
}

namespace tutu {
const char* (*g_ctx_error_hook)(const TutuCtx*) = nullptr;  // installed by tutu_b200.cu
}
extern "C" const char* tutu_last_error(const TutuCtx* ctx) {
  if (ctx && g_ctx_error_hook) return g_ctx_error_hook(ctx);
  return get_error().c_str();
}

extern "C" int tutu_traversal_tree_check(const TutuSceneDesc* desc, TutuTreeCheck* out) {
  if (!desc || !out) {
    set_error("tutu_traversal_tree_check: null argument");
    return TUTU_E_INVALID;
  }
  try {
    FlatScene fs;
    int rc = flatten_scene(desc, &fs);
    if (rc != TUTU_OK) return rc;
    build_wide_tree(&fs);
    memset(out, 0, sizeof(*out));
    out->n_leaves = (uint32_t)fs.leaf_box.size();
    out->binary_nodes = (uint32_t)fs.inner_fast.size();
    out->binary_depth = fs.depth_fast;
    out->wide_nodes = (uint32_t)fs.wide.size();
    out->wide_depth = fs.wide_depth;
    out->violations = verify_wide_tree(fs);
    uint64_t kids = 0;
    for (const WideNode& w : fs.wide) kids += (uint64_t)__builtin_popcount(w.imask) + (uint64_t)__builtin_popcount(w.lmask);
    out->wide_children = kids;
    return TUTU_OK;
  } catch (const std::bad_alloc&) {
    set_error("tutu_traversal_tree_check: out of host memory");
    return TUTU_E_NOMEM;
  } catch (const std::exception& e) {
    set_error(std::string("tutu_traversal_tree_check: ") + e.what());
    return TUTU_E_INVALID;
  }
}

extern "C" int tutu_bvh_build(const TutuPrim* prims, uint32_t n_prims, TutuBvhNode* nodes_out,
                              uint32_t* n_nodes_out) {
  if ((n_prims && !prims) || !nodes_out || !n_nodes_out) {
    set_error("tutu_bvh_build: null argument");
    return TUTU_E_INVALID;
  }
  for (uint32_t i = 0; i < n_prims; ++i)
    if (prims[i].type != TUTU_PRIM_TRIANGLE && prims[i].type != TUTU_PRIM_SPHERE) {
      set_error("tutu_bvh_build: unknown primitive type");
      return TUTU_E_INVALID;
    }
  try {
    std::vector<TutuBvhNode> nodes;
    build_bvh(prims, n_prims, &nodes);
    memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(TutuBvhNode));
    *n_nodes_out = (uint32_t)nodes.size();
    return TUTU_OK;
  } catch (const std::bad_alloc&) {
    set_error("tutu_bvh_build: out of host memory");
    return TUTU_E_NOMEM;
  } catch (const std::exception& e) {  // std::system_error from std::thread in the parallel build
    set_error(std::string("tutu_bvh_build: ") + e.what());
    return TUTU_E_INVALID;
  }
}

// ------------------------------------------------------------------------------------------
// scene files: "TUTUSCN1" header, then the arrays of TutuSceneDesc in declaration order
// ------------------------------------------------------------------------------------------
struct TutuSceneFile {
  TutuSceneDesc desc;
  std::vector<TutuPrim> prims;
  std::vector<TutuMaterial> materials;
  std::vector<TutuBvhNode> nodes;
  std::vector<TutuTexture> tex[4];
  std::vector<std::vector<float>> texdata[4];
};

namespace {
struct FileHeader {
  char magic[8];
  uint32_t version;
  uint32_t n_prims, n_materials, n_bvh_nodes;
  uint32_t n_tex[4];
  TutuCamera camera;
  float bkgcolor[3];
  float eta;
};
const char kMagic[8] = {'T', 'U', 'T', 'U', 'S', 'C', 'N', '1'};
struct FileCloser {
  void operator()(FILE* f) const {
    if (f) fclose(f);
  }
};
}  // namespace

extern "C" int tutu_scene_file_save(const TutuSceneDesc* desc, const char* path) {
  if (!desc || !path || desc->struct_size != sizeof(TutuSceneDesc)) {
    set_error("tutu_scene_file_save: bad argument");
    return TUTU_E_INVALID;
  }
  std::unique_ptr<FILE, FileCloser> f(fopen(path, "wb"));
  if (!f) {
    set_error(std::string("tutu_scene_file_save: cannot open ") + path);
    return TUTU_E_IO;
  }
  FileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, kMagic, 8);
  h.version = 1;
  h.n_prims = desc->n_prims;
  h.n_materials = desc->n_materials;
  h.n_bvh_nodes = desc->bvh_nodes ? desc->n_bvh_nodes : 0;
  for (int c = 0; c < 4; ++c) h.n_tex[c] = desc->n_tex[c];
  h.camera = desc->camera;
  memcpy(h.bkgcolor, desc->bkgcolor, 12);
  h.eta = desc->eta;
  bool ok = fwrite(&h, sizeof(h), 1, f.get()) == 1;
  if (h.n_prims) ok &= fwrite(desc->prims, sizeof(TutuPrim), h.n_prims, f.get()) == h.n_prims;
  if (h.n_materials)
    ok &= fwrite(desc->materials, sizeof(TutuMaterial), h.n_materials, f.get()) == h.n_materials;
  if (h.n_bvh_nodes)
    ok &= fwrite(desc->bvh_nodes, sizeof(TutuBvhNode), h.n_bvh_nodes, f.get()) == h.n_bvh_nodes;
  for (int c = 0; c < 4; ++c)
    for (uint32_t i = 0; i < h.n_tex[c]; ++i) {
      const TutuTexture& t = desc->tex[c][i];
      int32_t wh[2] = {t.width, t.height};
      ok &= fwrite(wh, sizeof(wh), 1, f.get()) == 1;
      size_t cnt = (size_t)t.width * t.height * 3;
      if (cnt) ok &= fwrite(t.rgb, sizeof(float), cnt, f.get()) == cnt;
    }
  if (!ok) {
    set_error(std::string("tutu_scene_file_save: short write to ") + path);
    return TUTU_E_IO;
  }
  return TUTU_OK;
}

extern "C" int tutu_scene_file_load(const char* path, TutuSceneFile** out) {
  if (!path || !out) {
    set_error("tutu_scene_file_load: null argument");
    return TUTU_E_INVALID;
  }
  *out = nullptr;
  std::unique_ptr<FILE, FileCloser> f(fopen(path, "rb"));
  if (!f) {
    set_error(std::string("tutu_scene_file_load: cannot open ") + path);
    return TUTU_E_IO;
  }
  FileHeader h;
  if (fread(&h, sizeof(h), 1, f.get()) != 1 || memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) {
    set_error(std::string("tutu_scene_file_load: not a TUTUSCN1 file: ") + path);
    return TUTU_E_IO;
  }
  // a corrupt header must not turn into a huge allocation: the counts are bounded by the file size
  long file_bytes = 0;
  {
    const long here = ftell(f.get());
    if (here < 0 || fseek(f.get(), 0, SEEK_END) != 0 || (file_bytes = ftell(f.get())) < 0 || fseek(f.get(), here, SEEK_SET) != 0) {
      set_error(std::string("tutu_scene_file_load: cannot size ") + path);
      return TUTU_E_IO;
    }
  }
  const uint64_t need = (uint64_t)h.n_prims * sizeof(TutuPrim) + (uint64_t)h.n_materials * sizeof(TutuMaterial) +
                        (uint64_t)h.n_bvh_nodes * sizeof(TutuBvhNode) +
                        ((uint64_t)h.n_tex[0] + h.n_tex[1] + h.n_tex[2] + h.n_tex[3]) * 8u;
  if (need + sizeof(FileHeader) > (uint64_t)file_bytes) {
    set_error(std::string("tutu_scene_file_load: truncated or corrupt file (header counts exceed the file size): ") + path);
    return TUTU_E_IO;
  }
  try {
  std::unique_ptr<TutuSceneFile> sf(new TutuSceneFile());
  bool ok = true;
  sf->prims.resize(h.n_prims);
  sf->materials.resize(h.n_materials);
  sf->nodes.resize(h.n_bvh_nodes);
  if (h.n_prims) ok &= fread(sf->prims.data(), sizeof(TutuPrim), h.n_prims, f.get()) == h.n_prims;
  if (ok && h.n_materials)
    ok &= fread(sf->materials.data(), sizeof(TutuMaterial), h.n_materials, f.get()) ==
          h.n_materials;
  if (ok && h.n_bvh_nodes)
    ok &= fread(sf->nodes.data(), sizeof(TutuBvhNode), h.n_bvh_nodes, f.get()) == h.n_bvh_nodes;
  for (int c = 0; ok && c < 4; ++c) {
    sf->tex[c].resize(h.n_tex[c]);
    sf->texdata[c].resize(h.n_tex[c]);
    for (uint32_t i = 0; ok && i < h.n_tex[c]; ++i) {
      int32_t wh[2];
      ok &= fread(wh, sizeof(wh), 1, f.get()) == 1;
      if (!ok || wh[0] < 0 || wh[1] < 0 || (uint64_t)wh[0] * (uint64_t)wh[1] * 12u > (uint64_t)file_bytes) {
        ok = false;
        break;
      }
      size_t cnt = (size_t)wh[0] * wh[1] * 3;
      sf->texdata[c][i].resize(cnt);
      if (cnt) ok &= fread(sf->texdata[c][i].data(), sizeof(float), cnt, f.get()) == cnt;
      sf->tex[c][i] = {wh[0], wh[1], sf->texdata[c][i].data()};
    }
  }
  if (!ok) {
    set_error(std::string("tutu_scene_file_load: truncated file: ") + path);
    return TUTU_E_IO;
  }
  TutuSceneDesc& d = sf->desc;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(TutuSceneDesc);
  d.n_prims = h.n_prims;
  d.prims = sf->prims.data();
  d.n_materials = h.n_materials;
  d.materials = sf->materials.data();
  d.n_bvh_nodes = h.n_bvh_nodes;
  d.bvh_nodes = h.n_bvh_nodes ? sf->nodes.data() : nullptr;
  for (int c = 0; c < 4; ++c) {
    d.n_tex[c] = h.n_tex[c];
    d.tex[c] = h.n_tex[c] ? sf->tex[c].data() : nullptr;
  }
  d.camera = h.camera;
  memcpy(d.bkgcolor, h.bkgcolor, 12);
  d.eta = h.eta;
  *out = sf.release();
  return TUTU_OK;
  } catch (const std::bad_alloc&) {
    set_error(std::string("tutu_scene_file_load: out of host memory reading ") + path);
    return TUTU_E_NOMEM;
  } catch (const std::exception& e) {
    set_error(std::string("tutu_scene_file_load: ") + e.what());
    return TUTU_E_IO;
  }
}

extern "C" const TutuSceneDesc* tutu_scene_file_desc(const TutuSceneFile* f) {
  return f ? &f->desc : nullptr;
}
extern "C" void tutu_scene_file_free(TutuSceneFile* f) { delete f; }

// ------------------------------------------------------------------------------------------
// synthetic workload (BASELINE.json configs[1]); libm-free so it is identical on every host
// ------------------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline float u01(uint64_t seed, uint64_t a, uint64_t b) {  // [0,1), 24 bits
  uint64_t h = splitmix64(splitmix64(seed ^ (a * 0xD1342543DE82EF95ull)) + b);
  return (float)(h >> 40) * (1.0f / 16777216.0f);
}
// parabolic stand-in for sin(2*pi*x), x in [0,1]
static inline float wave(float x) {
  return x < 0.5f ? 16.f * x * (0.5f - x) : -16.f * (x - 0.5f) * (1.f - x);
}

extern "C" int tutu_synth_heightfield(uint32_t G, uint64_t seed, TutuPrim* prims_out) {
  if (!prims_out || G == 0 || G > 16384) {
    set_error("tutu_synth_heightfield: bad argument");
    return TUTU_E_INVALID;
  }
  const float invG = 1.0f / (float)G;
  auto P = [&](uint32_t i, uint32_t j) -> V3 {
    float x = (float)i * invG, z = (float)j * invG;
    float h = 0.05f * u01(seed, i, j) + 0.2f * wave(x) * wave(z + 0.25f > 1.f ? z - 0.75f : z + 0.25f);
    return {x, h, z};
  };
  size_t k = 0;
  for (uint32_t i = 0; i < G; ++i)
    for (uint32_t j = 0; j < G; ++j) {
      V3 p00 = P(i, j), p01 = P(i, j + 1), p11 = P(i + 1, j + 1), p10 = P(i + 1, j);
      const V3 tri[2][3] = {{p00, p01, p11}, {p00, p11, p10}};
      for (int t = 0; t < 2; ++t) {
        TutuPrim& pr = prims_out[k++];
        memset(&pr, 0, sizeof(pr));
        pr.type = TUTU_PRIM_TRIANGLE;
        for (int c = 0; c < 3; ++c) st3(pr.v + 3 * c, tri[t][c]);
        V3 nn = normalized(cross(sub(tri[t][1], tri[t][0]), sub(tri[t][2], tri[t][0])));
        for (int c = 0; c < 3; ++c) {
          st3(pr.n + 3 * c, nn);
          pr.uv[2 * c + 0] = tri[t][c].x;
          pr.uv[2 * c + 1] = tri[t][c].z;
        }
        pr.material = 0;
        pr.tex_active = 0;
        pr.tex_diffuse = pr.tex_normal = pr.tex_roughness = pr.tex_metallic = -1;
      }
    }
  return TUTU_OK;
}

extern "C" int tutu_synth_rays(int kind, uint64_t seed, uint64_t first, uint64_t n_rays,
                               float* rays_out) {
  if (!rays_out || (kind != 0 && kind != 1)) {
    set_error("tutu_synth_rays: bad argument");
    return TUTU_E_INVALID;
  }
  for (uint64_t r = 0; r < n_rays; ++r) {
    const uint64_t id = first + r;
    float* o = rays_out + r * TUTU_RAY_FLOATS;
    V3 org, dir;
    float tmax;
    if (kind == 0) {
      org = {u01(seed, id, 0), 1.f + u01(seed, id, 1), u01(seed, id, 2)};
      V3 tgt = {u01(seed, id, 3), 0.f, u01(seed, id, 4)};
      dir = normalized(sub(tgt, org));
      tmax = 10.f;
    } else {
      org = {u01(seed, id, 0), -0.3f + 0.7f * u01(seed, id, 1), u01(seed, id, 2)};
      uint64_t k = 3;
      for (;;) {  // uniform direction by rejection from the unit ball
        V3 p = {2.f * u01(seed, id, k) - 1.f, 2.f * u01(seed, id, k + 1) - 1.f,
                2.f * u01(seed, id, k + 2) - 1.f};
        k += 3;
        float l2 = p.x * p.x + p.y * p.y + p.z * p.z;
        if (l2 <= 1.f && l2 > 1e-4f) {
          dir = normalized(p);
          break;
        }
      }
      tmax = 0.5f;
    }
    o[0] = org.x, o[1] = org.y, o[2] = org.z, o[3] = 0.f;
    o[4] = dir.x, o[5] = dir.y, o[6] = dir.z, o[7] = tmax;
  }
  return TUTU_OK;
}

// ------------------------------------------------------------------------------------------
// PPM output — PPMGenerator::writeHeader / writePixel file format (PPMGenerator.hpp:804-809, 840-842)
// ------------------------------------------------------------------------------------------
extern "C" int tutu_write_ppm(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8, int binary) {
  if (!path || (!rgb8 && (size_t)width * height != 0)) {
    tutu::set_error("tutu_write_ppm: null argument");
    return TUTU_E_INVALID;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    tutu::set_error(std::string("tutu_write_ppm: cannot open ") + path);
    return TUTU_E_IO;
  }
  const size_t npix = (size_t)width * height;
  bool ok = true;
  if (binary) {
    ok = fprintf(f, "P6\n%u %u\n255\n", width, height) > 0;
    ok = ok && fwrite(rgb8, 1, npix * 3, f) == npix * 3;
  } else {
    std::string buf;
    buf.reserve(npix * 12 + 32);
    buf += "P3\n" + std::to_string(width) + "\n" + std::to_string(height) + "\n255\n";
    char tmp[16];
    for (size_t i = 0; i < npix; ++i) {
      int n = snprintf(tmp, sizeof(tmp), "%d %d %d\n", (int)rgb8[3 * i], (int)rgb8[3 * i + 1], (int)rgb8[3 * i + 2]);
      buf.append(tmp, (size_t)n);
    }
    ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    tutu::set_error(std::string("tutu_write_ppm: write failed for ") + path);
    return TUTU_E_IO;
  }
  return TUTU_OK;
}
