// trace_variants.cuh — traversal flavours that were measured and lost: the warp walk with
// postponed leaves, persistent lanes with refill (shared- and local-memory stacks).  Compiled only into
// experiment builds (-DTUTU_EXPERIMENTS, tuturenderer_b200/build.py: build_variant); the shipped library
// does not contain them.  They return bit-identical hits (tools/gpu_variants.py compares them).
#pragma once
#include "trace.cuh"

namespace tutu {

// Warp-cooperative walk with postponed leaves (after Aila & Laine's speculative while-while): a lane
// that reaches a leaf parks it and keeps walking; parked leaves are tested by all their lanes
// together as soon as one lane cannot go on (it reached a second leaf or ran out of nodes).  All
// 32 lanes call this together; `valid` = the lane carries a ray.
constexpr int kNoLeaf = 0x7FFFFFFF;
template <bool ANY>
__device__ __forceinline__ bool traverse_warp(const DevScene& sc, const Ray& r, float dis, bool valid, Hit& best) {
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  bool done = !valid || !walk_begin(sc, w, r, dis);
  if (!valid) {
    w.best.t = FLT_MAX, w.best.u = 0.f, w.best.v = 0.f, w.best.slot = -1;
    w.regular = true;
  }
  int parked = kNoLeaf;
  bool out_of_nodes = false;  // stack empty, only the parked leaf is left
  const bool all_regular = __all_sync(0xFFFFFFFFu, done || w.regular);
  while (__any_sync(0xFFFFFFFFu, !done)) {
    // a lane is stuck when it holds a parked leaf and either stands on another leaf or has no node left
    const bool stuck = !done && parked != kNoLeaf && (out_of_nodes || w.cur < 0);
    if (__any_sync(0xFFFFFFFFu, stuck)) {
      if (!done && parked != kNoLeaf) {
        if (leaf_step<ANY>(sc, w, parked)) done = true;
        parked = kNoLeaf;
        if (out_of_nodes) done = true;
      }
      continue;
    }
    if (!done) {
      bool need_pop;
      if (w.cur >= 0) {
        need_pop = all_regular ? node_step<ANY, true>(sc, w, stack_ref, stack_t) : node_step<ANY, false>(sc, w, stack_ref, stack_t);
      } else {
        parked = w.cur;  // parked == kNoLeaf here, or the lane would be stuck
        need_pop = true;
      }
      if (need_pop && !walk_pop<ANY, 0>(sc, w, stack_ref, stack_t)) {
        if (parked != kNoLeaf)
          out_of_nodes = true;
        else
          done = true;
      }
    }
  }
  best = w.best;
  return best.slot >= 0;
}

// ---- persistent lanes with refill ----------------------------------------------------------------
// ncu on incoherent queues (profiles/r01_bdpt_full.txt: 6.5 of 32 lanes active per instruction in
// q_extend on the Veach room): a warp that walks a packet of 32 rays to completion idles behind its
// longest ray.  Here a warp owns a chunk of the queue (one global atomic per kRefillChunk rays) and
// hands idle lanes the next rays of the chunk whenever at least `refill_min` lanes are idle; between
// two such checks every lane runs kStepsPerCheck single steps (one node or one leaf, then a pop).
//   src(i, ray, dis) loads ray i;  sink(i, walk) stores its result (w.best).
constexpr unsigned kRefillChunk = 256;
constexpr int kStepsPerCheck = 8;

// one step of walk_shared's loop; true = the walk is over
template <bool ANY, bool REGULAR>
__device__ __forceinline__ bool walk_step_shared(const DevScene& sc, Walk& w, SharedStack<ANY>& st) {
  bool need_pop;
  if (w.cur >= 0) {
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
    need_pop = !(hl || hr);
    if (!need_pop) w.cur = swap ? k.y : k.x;
  } else {
    if (leaf_step<ANY>(sc, w, w.cur)) return true;
    need_pop = true;
  }
  if (need_pop) {
    for (;;) {
      if (st.sp == 0) return true;
      float t;
      st.pop(w.cur, t);
      if (!ANY && t > prune_limit<ANY>(sc, w)) continue;
      break;
    }
  }
  return false;
}

template <bool ANY, class Src, class Sink>
__device__ __forceinline__ void trace_refill(const DevScene& sc, unsigned long long n, unsigned long long* cursor,
                                             void* smem, Src src, Sink sink) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  Walk w;
  SharedStack<ANY> st;
  st.base = reinterpret_cast<typename SharedStack<ANY>::Word*>(smem) + threadIdx.x;
  st.stride = blockDim.x;
  bool active = false;
  unsigned long long my = 0;
  unsigned long long chunk_next = 0, chunk_end = 0;  // warp-uniform
  bool exhausted = false;                            // warp-uniform
  const int refill_min = sc.refill_min;
  for (;;) {
    const unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (!exhausted && (idle == 0xFFFFFFFFu || __popc(idle) >= refill_min)) {
      if (chunk_next == chunk_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kRefillChunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        chunk_next = base < n ? base : n;
        chunk_end = base + kRefillChunk < n ? base + kRefillChunk : n;
        if (chunk_next >= n) exhausted = true;
      }
      const unsigned long long left = chunk_end - chunk_next;
      const unsigned avail = left < 32ull ? (unsigned)left : 32u;
      const unsigned rank = __popc(idle & lt_mask);
      if (!active && rank < avail) {
        my = chunk_next + rank;
        Ray r;
        float dis;
        src(my, r, dis);
        st.sp = 0;
        if (walk_begin(sc, w, r, dis))
          active = true;
        else
          sink(my, w);  // missed the scene box: w.best is the miss record
      }
      const unsigned want = (unsigned)__popc(idle);
      chunk_next += want < avail ? want : avail;
    }
    if (__ballot_sync(0xFFFFFFFFu, active) == 0u) {
      if (exhausted) break;
      continue;
    }
#pragma unroll 1
    for (int k = 0; k < kStepsPerCheck; ++k) {
      if (active) {
        const bool over = w.regular ? walk_step_shared<ANY, true>(sc, w, st) : walk_step_shared<ANY, false>(sc, w, st);
        if (over) {
          sink(my, w);
          active = false;
        }
      }
    }
  }
}

// Persistent-thread tracer: every warp owns a chunk of the ray queue (one global atomicAdd per
// kChunk rays); between rounds the lanes whose ray is finished are handed the next rays of the
// chunk (ballot + popc ranks, no further atomics), so a warp keeps its lanes busy instead of
// idling behind its longest ray.
//   src(i, ray, dis)  loads ray i;  sink(i, walk)  stores its result.
constexpr unsigned kChunk = 256;

template <bool ANY, class Src, class Sink>
__device__ __forceinline__ void trace_persistent(const DevScene& sc, unsigned long long n,
                                                 unsigned long long* cursor, Src src, Sink sink) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  bool active = false;
  unsigned long long my = 0;
  unsigned long long chunk_next = 0, chunk_end = 0;  // warp-uniform
  bool exhausted = false;                            // warp-uniform
  for (;;) {
    const unsigned need = __ballot_sync(0xFFFFFFFFu, !active);
    if (!exhausted && (need == 0xFFFFFFFFu || __popc(need) >= sc.refill_min)) {
      if (chunk_next == chunk_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)kChunk);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        chunk_next = base < n ? base : n;
        chunk_end = base + kChunk < n ? base + kChunk : n;
        if (chunk_next >= n) {
          exhausted = true;
          chunk_end = chunk_next;
        }
      }
      const unsigned long long avail = chunk_end - chunk_next;
      const unsigned rank = __popc(need & lt_mask);
      if (!active && rank < avail) {
        my = chunk_next + rank;
        Ray r;
        float dis;
        src(my, r, dis);
        if (walk_begin(sc, w, r, dis))
          active = true;
        else
          sink(my, w);  // miss without entering the tree
      }
      const unsigned cnt = (unsigned)__popc(need);
      chunk_next += cnt < avail ? cnt : avail;
    }
    const unsigned act = __ballot_sync(0xFFFFFFFFu, active);
    if (act == 0u) {
      if (exhausted) break;
      continue;
    }
    // the exact (NaN-literal) slab test is valid for every ray; the FMNMX flavour only for regular ones
    const bool all_regular = __all_sync(0xFFFFFFFFu, !active || w.regular);
    if (active) {
      bool more;
      if (all_regular)
        more = walk_round<ANY, 0, false, true>(sc, w, stack_ref, stack_t, nullptr);
      else
        more = walk_round<ANY, 0, false, false>(sc, w, stack_ref, stack_t, nullptr);
      if (!more) {
        sink(my, w);
        active = false;
      }
    }
  }
}

// VARIANT 0: LOOP + exact slab test; 1: LOOP + FMNMX slab test for regular rays;
// 3: structured walk (walk_structured) + FMNMX slab test for regular rays;
// 2: rounds (inner-node phase / leaf phase) + FMNMX for regular rays.
template <bool ANY, int VARIANT>
__device__ __forceinline__ bool traverse_variant(const DevScene& sc, const Ray& r, float dis, Hit& best) {
  Walk w;
  int stack_ref[kStackSize];
  float stack_t[kStackSize];
  if (walk_begin(sc, w, r, dis)) {
    if (VARIANT == 0) {
      walk_loop<ANY, 0, false>(sc, w, stack_ref, stack_t);
    } else if (VARIANT == 1) {
      if (w.regular)
        walk_loop<ANY, 0, true>(sc, w, stack_ref, stack_t);
      else
        walk_loop<ANY, 0, false>(sc, w, stack_ref, stack_t);
    } else if (VARIANT == 3) {
      if (w.regular)
        walk_structured<ANY, true>(sc, w, stack_ref, stack_t);
      else
        walk_structured<ANY, false>(sc, w, stack_ref, stack_t);
    } else {
      if (w.regular) {
        while (walk_round<ANY, 0, false, true>(sc, w, stack_ref, stack_t, nullptr)) {
        }
      } else {
        while (walk_round<ANY, 0, false, false>(sc, w, stack_ref, stack_t, nullptr)) {
        }
      }
    }
  }
  best = w.best;
  return best.slot >= 0;
}

}  // namespace tutu
