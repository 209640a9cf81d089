// Host side of libtutu_b200: the compressed 8-wide traversal tree (tutu_internal.hpp: WideNode).
//
// For a regular ray (every 1/d finite) the reference's answer depends only on the LEAF boxes and the
// primitives (host_scene.cpp, "traversal tree for regular rays"): inner boxes are pure culling aids, and any
// inner box that CONTAINS the boxes below it keeps "leaf box hit => every ancestor box hit", because the slab
// test fl(fl(plane - o) * inv) is monotone in `plane`.  That allows (a) the binned-SAH topology of
// build_fast_tree and, here, (b) collapsing it to 8 children per node with the child boxes quantised OUTWARDS to
// 8 bits per plane (after Ylitie, Karras, Laine: "Efficient incoherent ray traversal on GPUs through compressed
// wide BVHs", HPG 2017): a third of the node visits of the binary tree and less than half its bytes per ray.
//
// What is different from that paper is the arithmetic contract.  The device does not fold the ray into the
// quantisation frame; it DECODES every plane to the fp32 value
//       dec(q) = fma(as_float(0x4B000000 | q), scale, base2)           (one correctly rounded operation)
// and then runs the reference's own fl(fl(plane - o) * inv) on it.  This file picks q with the very same
// fmaf(): q_lo = the largest q with dec(q) <= the child's exact lower plane, q_hi = the smallest q with
// dec(q) >= its exact upper plane.  Containment of the decoded box is therefore checked value by value on the
// host, not argued from error bounds, and monotonicity of the slab test does the rest.  The exact leaf boxes
// (WideLeafBox) are tested on the device whenever a primitive test accepts, so the hit set is the reference's.
//
// Pure C++; built with -ffp-contract=off like host_scene.cpp (fmaf is explicit where a fused operation is meant).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "tutu_internal.hpp"

namespace tutu {
namespace {

inline float bits_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline float wide_decode(uint32_t q, float scale, float base2) { return fmaf(bits_float(0x4B000000u | q), scale, base2); }

inline float box_area(const Box& b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return 2.f * (dx * dy + dy * dz + dz * dx);
}
inline Box box_join(const Box& a, const Box& b) {
  Box r;
  for (int k = 0; k < 3; ++k) r.lo[k] = fminf(a.lo[k], b.lo[k]), r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
  return r;
}
inline Box child_box(const InnerNode& n, int right) {
  Box b;
  for (int a = 0; a < 3; ++a) b.lo[a] = n.box[6 * right + 2 * a], b.hi[a] = n.box[6 * right + 2 * a + 1];
  return b;
}

struct Child {
  int32_t ref;  // binary-tree ref: >= 0 inner node, < 0 ~(slot | sphere bit)
  Box box;
};

// Quantisation frame of one axis: scale = 2^e and base2 with dec(0) <= lo and dec(255) >= hi, both verified.
bool pick_frame(float lo, float hi, float* scale_out, float* base2_out) {
  if (!std::isfinite(lo) || !std::isfinite(hi) || hi < lo) return false;
  const float extent = hi - lo;
  if (!std::isfinite(extent)) return false;
  int e;
  if (extent > 0.f) {
    frexpf(extent / 255.f, &e);  // extent / 255 = m * 2^e, m in [0.5, 1): 2^e >= extent / 255
  } else {
    int el;
    frexpf(lo == 0.f ? 1.f : fabsf(lo), &el);
    e = el - 30;  // far below the resolution of `lo`: every q decodes to lo
  }
  if (e < -120) e = -120;
  for (; e <= 100; ++e) {
    const float scale = ldexpf(1.f, e);
    float base2 = lo - 8388608.f * scale;
    if (!std::isfinite(base2)) return false;
    int steps = 0;
    while (wide_decode(0, scale, base2) > lo && steps < 8) base2 = nextafterf(base2, -FLT_MAX), ++steps;
    if (wide_decode(0, scale, base2) > lo) continue;
    if (wide_decode(255, scale, base2) < hi) continue;
    *scale_out = scale;
    *base2_out = base2;
    return true;
  }
  return false;
}

uint8_t quant_lo(float plane, float scale, float base2) {  // largest q with dec(q) <= plane (dec(0) <= plane holds)
  int lo = 0, hi = 255;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (wide_decode((uint32_t)mid, scale, base2) <= plane)
      lo = mid;
    else
      hi = mid - 1;
  }
  return (uint8_t)lo;
}
uint8_t quant_hi(float plane, float scale, float base2) {  // smallest q with dec(q) >= plane (dec(255) >= plane holds)
  int lo = 0, hi = 255;
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (wide_decode((uint32_t)mid, scale, base2) >= plane)
      hi = mid;
    else
      lo = mid + 1;
  }
  return (uint8_t)lo;
}

}  // namespace

// Collapses fs->inner_fast (root fs->root_ref_fast) into fs->wide / wleaf / wbox.  On any failure (non-finite
// planes, frames that cannot be verified) the wide arrays stay empty and the device keeps the binary walk.
void build_wide_tree(FlatScene* fs) {
  fs->wide.clear();
  fs->wleaf.clear();
  fs->wbox.clear();
  fs->wide_depth = 0;
  const uint32_t n = (uint32_t)fs->leaf_box.size();
  if (fs->empty || n == 0) return;
  for (const Box& b : fs->leaf_box)
    for (int a = 0; a < 3; ++a)
      if (!std::isfinite(b.lo[a]) || !std::isfinite(b.hi[a])) return;
  const PodVec<InnerNode>& bin = fs->inner_fast;

  // ---- which binary nodes become wide nodes: the surface-area dynamic programme of Ylitie et al. (section 3.1) ----
  // C(m, i) = least sum of wide-node surface areas with which the subtree of binary node m can be represented
  // as at most i roots (i = 1: one wide node at m; leaves are single primitives and cost the same in any tree):
  //   C(m, 1) = A(m) + min_k C(left, k) + C(right, 8 - k)          (m's wide node has <= 8 child slots)
  //   C(m, i) = min(C(m, i - 1), min_k C(left, k) + C(right, i - k))
  // The table is filled children first: reverse pre-order of the tree (the host SAH tree is stored in pre-order,
  // a device-built Karras tree is not, so the order is computed).  split[m][i-1] = k chosen for i roots
  // (0 = "use the i - 1 solution").
  const size_t nb_nodes = bin.size();
  std::vector<float> cost(nb_nodes * 8);
  std::vector<uint8_t> split(nb_nodes * 8);
  std::vector<uint32_t> preorder;
  preorder.reserve(nb_nodes);
  if (fs->root_ref_fast >= 0) {
    std::vector<uint32_t> todo;
    todo.push_back((uint32_t)fs->root_ref_fast);
    while (!todo.empty()) {
      const uint32_t m = todo.back();
      todo.pop_back();
      if (m >= nb_nodes || preorder.size() >= nb_nodes) return;  // not a tree
      preorder.push_back(m);
      if (bin[m].right >= 0) todo.push_back((uint32_t)bin[m].right);
      if (bin[m].left >= 0) todo.push_back((uint32_t)bin[m].left);
    }
  }
  auto C = [&](int32_t ref, int i) -> float { return ref < 0 ? 0.f : cost[(size_t)ref * 8 + (size_t)(i - 1)]; };
  for (size_t pi = preorder.size(); pi-- > 0;) {
    const size_t m = preorder[pi];
    const InnerNode& b = bin[m];
    const float area = box_area(box_join(child_box(b, 0), child_box(b, 1)));
    auto distribute = [&](int slots, uint8_t* k_out) {
      float best = FLT_MAX;
      uint8_t bk = 1;
      for (int k = 1; k < slots; ++k) {
        const float c = C(b.left, k) + C(b.right, slots - k);
        if (c < best) best = c, bk = (uint8_t)k;
      }
      *k_out = bk;
      return best;
    };
    uint8_t k8;
    cost[m * 8 + 0] = area + distribute(8, &k8);
    split[m * 8 + 0] = k8;
    for (int i = 2; i <= 8; ++i) {
      uint8_t k;
      const float d = distribute(i, &k);
      if (d < cost[m * 8 + (size_t)(i - 2)]) {
        cost[m * 8 + (size_t)(i - 1)] = d;
        split[m * 8 + (size_t)(i - 1)] = k;
      } else {
        cost[m * 8 + (size_t)(i - 1)] = cost[m * 8 + (size_t)(i - 2)];
        split[m * 8 + (size_t)(i - 1)] = 0;
      }
    }
  }

  struct Work {
    int32_t ref;      // binary subtree this wide node stands for (>= 0), or a single leaf ref (< 0, only the root)
    uint32_t depth;
  };
  std::vector<Work> work;  // work[i] describes wide node i (breadth first: children of a node are contiguous)
  work.push_back({fs->root_ref_fast, 1});
  std::vector<WideNode> out;
  std::vector<uint32_t> leaf_order;  // wide-leaf index -> leaf code (slot | sphere bit)
  out.reserve(n / 4 + 1);
  leaf_order.reserve(n);

  // the roots that represent subtree `ref` with at most i roots, appended to kids[]
  struct Pending {
    int32_t ref;
    int i;
    Box box;
  };
  for (size_t wi = 0; wi < work.size(); ++wi) {
    const Work w = work[wi];
    fs->wide_depth = std::max(fs->wide_depth, w.depth);
    Child kids[8];
    int nk = 0;
    if (w.ref < 0) {
      kids[nk++] = {w.ref, fs->leaf_box[(~(uint32_t)w.ref) & SLOT_MASK]};
    } else {
      const InnerNode& b = bin[(size_t)w.ref];
      const int k = split[(size_t)w.ref * 8 + 0];
      Pending st[16];
      int sp = 0;
      st[sp++] = {b.right, 8 - k, child_box(b, 1)};
      st[sp++] = {b.left, k, child_box(b, 0)};
      while (sp) {
        const Pending p = st[--sp];
        if (p.ref < 0 || p.i == 1) {
          kids[nk++] = {p.ref, p.box};
          continue;
        }
        int i = p.i;
        while (i > 1 && split[(size_t)p.ref * 8 + (size_t)(i - 1)] == 0) --i;  // "use the i - 1 solution"
        if (i == 1) {
          kids[nk++] = {p.ref, p.box};
          continue;
        }
        const InnerNode& c = bin[(size_t)p.ref];
        const int kk = split[(size_t)p.ref * 8 + (size_t)(i - 1)];
        st[sp++] = {c.right, i - kk, child_box(c, 1)};
        st[sp++] = {c.left, kk, child_box(c, 0)};
      }
    }
    Box nb = kids[0].box;
    for (int k = 1; k < nk; ++k) nb = box_join(nb, kids[k].box);

    // slot assignment: slot bit a set = the child lies on the + side of axis a.  A ray travelling towards +a
    // visits the slots with bit a clear first (wide.cuh: priority = slot ^ octant), i.e. roughly front to back.
    int slot_of[8], kid_in_slot[8];
    for (int s = 0; s < 8; ++s) kid_in_slot[s] = -1;
    {
      float cost[8][8];
      for (int k = 0; k < nk; ++k) {
        float off[3];
        for (int a = 0; a < 3; ++a)
          off[a] = (0.5f * kids[k].box.lo[a] + 0.5f * kids[k].box.hi[a]) - (0.5f * nb.lo[a] + 0.5f * nb.hi[a]);
        for (int s = 0; s < 8; ++s) cost[k][s] = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
        slot_of[k] = -1;
      }
      for (int round = 0; round < nk; ++round) {
        int bk = -1, bs = -1;
        float bc = -FLT_MAX;
        for (int k = 0; k < nk; ++k) {
          if (slot_of[k] >= 0) continue;
          for (int s = 0; s < 8; ++s)
            if (kid_in_slot[s] < 0 && (bk < 0 || cost[k][s] > bc)) bc = cost[k][s], bk = k, bs = s;
        }
        slot_of[bk] = bs;
        kid_in_slot[bs] = bk;
      }
    }

    WideNode node;
    memset(&node, 0, sizeof(node));
    for (int a = 0; a < 3; ++a)
      if (!pick_frame(nb.lo[a], nb.hi[a], &node.scale[a], &node.base2[a])) {
        fs->wide_depth = 0;
        return;
      }
    node.child_base = (uint32_t)work.size();
    node.leaf_base = (uint32_t)leaf_order.size();
    for (int s = 0; s < 8; ++s) {
      const int k = kid_in_slot[s];
      if (k < 0) {  // empty slot: inverted box, never hit
        for (int a = 0; a < 3; ++a) node.qlo[a][s] = 255, node.qhi[a][s] = 0;
        continue;
      }
      for (int a = 0; a < 3; ++a) {
        node.qlo[a][s] = quant_lo(kids[k].box.lo[a], node.scale[a], node.base2[a]);
        node.qhi[a][s] = quant_hi(kids[k].box.hi[a], node.scale[a], node.base2[a]);
      }
      if (kids[k].ref >= 0) {
        node.imask |= (uint8_t)(1u << s);
        work.push_back({kids[k].ref, w.depth + 1});
      } else {
        node.lmask |= (uint8_t)(1u << s);
        leaf_order.push_back(~(uint32_t)kids[k].ref);
      }
    }
    out.push_back(node);
  }
  if (leaf_order.size() != n) {  // cannot happen: every leaf of the binary tree is reached exactly once
    fs->wide_depth = 0;
    return;
  }
  fs->wleaf.resize(n);
  fs->wbox.resize(n);
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t code = leaf_order[k], slot = code & SLOT_MASK;
    WideLeaf& L = fs->wleaf[k];
    memcpy(L.f, fs->geom[slot].f, sizeof(L.f));
    L.code = code;
    L.pad[0] = L.pad[1] = L.pad[2] = 0;
    WideLeafBox& B = fs->wbox[k];
    for (int a = 0; a < 3; ++a) B.lo[a] = fs->leaf_box[slot].lo[a], B.hi[a] = fs->leaf_box[slot].hi[a];
    B.pad[0] = B.pad[1] = 0.f;
  }
  fs->wide = std::move(out);
}

// Host-side check used by the CPU tests: every child box of every wide node, decoded exactly as the device
// decodes it, contains the exact boxes of all leaves below it; every leaf is referenced once.  Returns the
// number of violations.
uint64_t verify_wide_tree(const FlatScene& fs) {
  if (fs.wide.empty()) return 0;
  uint64_t bad = 0;
  std::vector<uint8_t> seen(fs.wleaf.size(), 0);
  struct Item {
    uint32_t node;
    Box bound;  // decoded box this subtree must stay inside (root: unbounded)
  };
  std::vector<Item> stack;
  Box all;
  for (int a = 0; a < 3; ++a) all.lo[a] = -FLT_MAX, all.hi[a] = FLT_MAX;
  stack.push_back({0, all});
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    const WideNode& nd = fs.wide[it.node];
    uint32_t inner_seen = 0, leaf_seen = 0;
    for (int s = 0; s < 8; ++s) {
      const bool inner = (nd.imask >> s) & 1, leaf = (nd.lmask >> s) & 1;
      if (inner && leaf) ++bad;
      if (!inner && !leaf) {
        if (!(nd.qlo[0][s] > nd.qhi[0][s])) ++bad;  // empty slots must be inverted
        continue;
      }
      Box d;
      for (int a = 0; a < 3; ++a) {
        d.lo[a] = wide_decode(nd.qlo[a][s], nd.scale[a], nd.base2[a]);
        d.hi[a] = wide_decode(nd.qhi[a][s], nd.scale[a], nd.base2[a]);
      }
      if (inner) {
        // containment is transitive only through EXACT boxes: a child's decoded box must contain the leaves
        // below it, which is checked at the leaves against every ancestor's decoded box via `bound` below
        Box nb;
        for (int a = 0; a < 3; ++a) nb.lo[a] = fmaxf(d.lo[a], it.bound.lo[a]), nb.hi[a] = fminf(d.hi[a], it.bound.hi[a]);
        stack.push_back({nd.child_base + inner_seen, nb});
        ++inner_seen;
      } else {
        const uint32_t k = nd.leaf_base + leaf_seen;
        ++leaf_seen;
        if (k >= fs.wleaf.size() || seen[k]) {
          ++bad;
          continue;
        }
        seen[k] = 1;
        const WideLeafBox& lb = fs.wbox[k];
        for (int a = 0; a < 3; ++a) {
          if (!(d.lo[a] <= lb.lo[a]) || !(d.hi[a] >= lb.hi[a])) ++bad;
          if (!(it.bound.lo[a] <= lb.lo[a]) || !(it.bound.hi[a] >= lb.hi[a])) ++bad;
        }
      }
    }
  }
  for (uint8_t s : seen)
    if (!s) ++bad;
  return bad;
}

}  // namespace tutu
