// landmass_b200 — pieces shared by the A* kernels (k_astar.cu, k_astar_quad.cu): the exact
// emulation of Rust's BinaryHeap<Reverse<NodeRef>> (astar.rs:56-83),
// edge-record loads and the per-query phase codes.
#pragma once
#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {
namespace astar_detail {

constexpr float F_INF = __builtin_huge_valf();
constexpr uint32_t FULL = 0xffffffffu;

struct HeapEntry {  // 16 bytes: one 128-bit shared-memory access, half a sector when spilled
  float cost, estimate;
  uint32_t index;
  uint32_t state;
};

// NodeRef::partial_cmp (astar.rs:67-77) as `a <= b`
__device__ __forceinline__ bool noderef_le(const HeapEntry& a, const HeapEntry& b) {
  // Less, or Equal-estimates with Reverse(a.cost) <= Reverse(b.estimate); unordered (NaN) -> false
  return (a.estimate < b.estimate) | ((a.estimate == b.estimate) & (b.estimate <= a.cost));
}
// Reverse<T>::le(x, y) = y.0 <= x.0
__device__ __forceinline__ bool rev_le(const HeapEntry& x, const HeapEntry& y) { return noderef_le(y, x); }

// Open list.  The SP = false members touch shared memory only (valid while every index they
// can reach is < HS); SP = true adds the in-place spill to the slot's arena.  The
// caller picks per operation from the current length, so the common case carries no
// per-access branch.
template <uint32_t STRIDE, uint32_t HS>  // HS: entries held in shared memory (even)
struct Heap {
  static_assert(HS % 2 == 0, "HS must be even: spilled sibling pairs (2k+1, 2k+2) then share a 32-byte sector");
  uint4* sm;      // this query's column: entry i at sm[i * STRIDE]
  uint4* gm;      // spill: entry i (>= HS) at gm[i - HS + 1]
  uint32_t len;
  template <bool SP>
  __device__ __forceinline__ HeapEntry get(uint32_t i) const {
    uint4 v;
    if (!SP || i < HS)
      v = sm[i * STRIDE];
    else
      v = gm[i - HS + 1];
    return {__uint_as_float(v.x), __uint_as_float(v.y), v.z, v.w};
  }
  template <bool SP>
  __device__ __forceinline__ void set(uint32_t i, const HeapEntry& e) {
    uint4 v = make_uint4(__float_as_uint(e.cost), __float_as_uint(e.estimate), e.index, e.state);
    if (!SP || i < HS)
      sm[i * STRIDE] = v;
    else
      gm[i - HS + 1] = v;
  }
  template <bool SP>
  __device__ __forceinline__ void sift_up(uint32_t pos, const HeapEntry& hole) {
    while (pos > 0) {
      uint32_t parent = (pos - 1) >> 1;
      HeapEntry p = get<SP>(parent);
      if (rev_le(hole, p)) break;
      set<SP>(pos, p);
      pos = parent;
    }
    set<SP>(pos, hole);
  }
  template <bool SP>
  __device__ __forceinline__ void push_t(const HeapEntry& e) {
    uint32_t old_len = len++;
    sift_up<SP>(old_len, e);
  }
  __device__ __forceinline__ void push(const HeapEntry& e) {
    if (len < HS)
      push_t<false>(e);
    else
      push_t<true>(e);
  }
  template <bool SP>
  __device__ __forceinline__ HeapEntry pop_t() {
    HeapEntry item = get<SP>(--len);
    if (len > 0) {
      HeapEntry top = get<SP>(0);
      HeapEntry hole = item;  // swap(item, data[0]); the hole element is the old last
      item = top;
      uint32_t end = len, pos = 0, child = 1;
      while (child + 1 < end) {  // child <= end - 2
        HeapEntry a = get<SP>(child), b = get<SP>(child + 1);
        bool right = rev_le(a, b);
        set<SP>(pos, right ? b : a);
        pos = child + (right ? 1u : 0u);
        child = 2 * pos + 1;
      }
      if (child + 1 == end) {
        set<SP>(pos, get<SP>(child));
        pos = child;
      }
      sift_up<SP>(pos, hole);
    }
    return item;
  }
  __device__ __forceinline__ HeapEntry pop() {
    if (len <= HS) return pop_t<false>();
    return pop_t<true>();
  }
};

template <class T>
__device__ __forceinline__ T sel4(uint32_t j, T a0, T a1, T a2, T a3) {
  return j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));
}

__device__ __forceinline__ EdgeRec ld_rec(const EdgeRec* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  EdgeRec r;
  r.mx = __uint_as_float(a.x);
  r.my = __uint_as_float(a.y);
  r.mz = __uint_as_float(a.z);
  r.twin = a.w;
  r.poly = b.x;
  r.first_edge = b.y;
  r.info = b.z;
  r.aux = b.w;
  return r;
}

// ---------------------------------------------------------------------------------------------
// `best_estimates` (HashMap<PathNode, f32> in the reference, astar.rs:131), two layouts per slot:
//  * dense   — f32 best[n_states], +inf = absent.
//  * compact — padded ids only.  One 8-byte entry per POLYGON {estimate, local edge of the one
//    state tracked} (a search enters most polygons through a single edge: always in tree-like
//    meshes), an f32 per special state (links, Start, End) and a small open-addressing overflow
//    table {state, estimate} for further states of an already tracked polygon.  Half the bytes
//    of the dense array (so more queries fit in HBM) and four polygons per 32-byte sector instead
//    of two.  A query that fills the overflow table is re-run with the dense layout.
//  * hashed  — for meshes so large that a search touches a small fraction of the states (1M
//    polygons: ~1 %): an open-addressing table {state, estimate} (linear probing, power-of-two
//    capacity >= 2 x node capacity, so at most half full: a query that would exceed that is re-run
//    with a doubled arena).  16 bytes per touched state instead of 4 per existing state.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t BP_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t BP_SPECIAL = 0xFFFFFFFEu;

template <int LAYOUT>  // 0 dense, 1 compact, 2 hashed, 3 none (forest meshes)
struct BestStore;

template <>
struct BestStore<0> {
  typedef float Raw;
  float* b;
  __device__ __forceinline__ void init(uint8_t* slot_base, const AStarArena&) { b = reinterpret_cast<float*>(slot_base); }
  __device__ __forceinline__ Raw probe(uint32_t st) const { return __ldcg(b + st); }
  __device__ __forceinline__ Raw probe_edge(uint32_t st) const { return __ldcg(b + st); }
  typedef Raw RootRaw;
  __device__ __forceinline__ RootRaw probe_root(uint32_t st) const { return probe(st); }
  __device__ __forceinline__ float root_value(uint32_t, RootRaw r) const { return r; }
  __device__ __forceinline__ float value(uint32_t, Raw r) const { return r; }
  __device__ __forceinline__ float value_edge(uint32_t, Raw r) const { return r; }
  __device__ __forceinline__ float value_fast(uint32_t, Raw r, bool& need_table) const {
    need_table = false;
    return r;
  }
  __device__ __forceinline__ bool store(uint32_t st, Raw, float est) {
    b[st] = est;
    return true;
  }
};

template <>
struct BestStore<1> {
  typedef uint2 Raw;
  uint2* bp;    // [n_polys] {estimate bits, tracked local edge | BP_EMPTY}
  float* sp;    // [n_links + 2]
  uint2* ovf;   // [ovf_cap] {state | BP_EMPTY, estimate bits}
  uint32_t n_ids, kshift, kmask, ovf_mask, ovf_count;
  __device__ __forceinline__ void init(uint8_t* slot_base, const AStarArena& a) {
    bp = reinterpret_cast<uint2*>(slot_base);
    sp = reinterpret_cast<float*>(slot_base + a.special_offset);
    ovf = reinterpret_cast<uint2*>(slot_base + a.ovf_offset);
    n_ids = a.n_ids;
    kshift = a.kshift;
    kmask = (1u << a.kshift) - 1u;
    ovf_mask = a.ovf_cap - 1u;
    ovf_count = 0;
  }
  __device__ __forceinline__ Raw probe(uint32_t st) const {
    if (st < n_ids) return __ldcg(bp + (st >> kshift));
    return make_uint2(__float_as_uint(__ldcg(sp + (st - n_ids))), BP_SPECIAL);
  }
  __device__ __forceinline__ Raw probe_edge(uint32_t st) const { return __ldcg(bp + (st >> kshift)); }
  // Probe of the open list's root, issued before the pop and consumed after it.  Two loads into
  // separate registers and no write to either before the consumer: `probe()` patches .y of a
  // special state's result right behind the (predicated-off) 64-bit load of the same register
  // pair, and that write waits for the load — the whole memory latency, in front of the pop
  // it was issued early to overlap with (ncu: 13.8 % of the kernel's stall samples).
  struct RootRaw {
    uint2 e;
    float special;
  };
  __device__ __forceinline__ RootRaw probe_root(uint32_t st) const {
    const bool edge = st < n_ids;
    RootRaw r;
    r.e = __ldcg(bp + (edge ? st >> kshift : 0u));
    r.special = 0.0f;
    if (!edge) r.special = __ldcg(sp + (st - n_ids));
    return r;
  }
  __device__ __forceinline__ float root_value(uint32_t st, const RootRaw& r) const {
    if (st >= n_ids) return r.special;
    return value_edge(st, r.e);
  }
  __device__ __forceinline__ uint32_t ovf_slot(uint32_t st) const { return (st * 2654435761u) >> 7; }
  __device__ float ovf_get(uint32_t st) const {
    for (uint32_t h = ovf_slot(st);; h++) {
      const uint2 e = __ldcg(ovf + (h & ovf_mask));
      if (e.x == st) return __uint_as_float(e.y);
      if (e.x == BP_EMPTY) return __builtin_huge_valf();
    }
  }
  __device__ bool ovf_set(uint32_t st, float est) {
    for (uint32_t h = ovf_slot(st);; h++) {
      const uint2 e = __ldcg(ovf + (h & ovf_mask));
      if (e.x == st || e.x == BP_EMPTY) {
        if (e.x == BP_EMPTY && ++ovf_count > (ovf_mask + 1u) / 2u) return false;  // too full: go dense
        ovf[h & ovf_mask] = make_uint2(st, __float_as_uint(est));
        return true;
      }
    }
  }
  __device__ __forceinline__ float value(uint32_t st, Raw r) const {
    if (r.y == BP_SPECIAL || r.y == (st & kmask)) return __uint_as_float(r.x);
    if (r.y == BP_EMPTY) return __builtin_huge_valf();
    return ovf_get(st);
  }
  // edge states only: branch-free for "tracked" and "absent"; anything else is in the table
  __device__ __forceinline__ float value_fast(uint32_t st, Raw r, bool& need_table) const {
    const bool tracked = r.y == (st & kmask);
    need_table = !tracked & (r.y != BP_EMPTY);
    return tracked ? __uint_as_float(r.x) : __builtin_huge_valf();
  }
  // edge states only (never BP_SPECIAL)
  __device__ __forceinline__ float value_edge(uint32_t st, Raw r) const {
    if (r.y == (st & kmask)) return __uint_as_float(r.x);
    if (r.y == BP_EMPTY) return __builtin_huge_valf();
    return ovf_get(st);
  }
  __device__ __forceinline__ bool store(uint32_t st, Raw r, float est) {
    if (r.y == BP_SPECIAL) {
      sp[st - n_ids] = est;
    } else if (r.y == BP_EMPTY || r.y == (st & kmask)) {
      bp[st >> kshift] = make_uint2(__float_as_uint(est), st & kmask);
    } else {
      return ovf_set(st, est);
    }
    return true;
  }
};

template <>
struct BestStore<2> {
  typedef uint2 Raw;  // the entry at the state's home slot {key, estimate bits}
  uint2* tab;
  uint32_t mask, count, limit;
  __device__ __forceinline__ void init(uint8_t* slot_base, const AStarArena& a) {
    tab = reinterpret_cast<uint2*>(slot_base);
    mask = a.hash_cap - 1u;
    count = 0;
    limit = a.hash_cap / 2u;
  }
  __device__ __forceinline__ uint32_t home(uint32_t st) const { return (st * 2654435761u) >> 4; }
  __device__ __forceinline__ Raw probe(uint32_t st) const { return __ldcg(tab + (home(st) & mask)); }
  __device__ __forceinline__ Raw probe_edge(uint32_t st) const { return probe(st); }
  typedef Raw RootRaw;
  __device__ __forceinline__ RootRaw probe_root(uint32_t st) const { return probe(st); }
  __device__ __forceinline__ float root_value(uint32_t st, RootRaw r) const { return value(st, r); }
  __device__ float find(uint32_t st) const {
    for (uint32_t h = home(st) + 1;; h++) {
      const uint2 e = __ldcg(tab + (h & mask));
      if (e.x == st) return __uint_as_float(e.y);
      if (e.x == BP_EMPTY) return __builtin_huge_valf();
    }
  }
  __device__ __forceinline__ float value_fast(uint32_t st, Raw r, bool& need_table) const {
    const bool here = r.x == st;
    need_table = !here & (r.x != BP_EMPTY);
    return here ? __uint_as_float(r.y) : __builtin_huge_valf();
  }
  __device__ __forceinline__ float value(uint32_t st, Raw r) const {
    if (r.x == st) return __uint_as_float(r.y);
    if (r.x == BP_EMPTY) return __builtin_huge_valf();
    return find(st);
  }
  __device__ __forceinline__ float value_edge(uint32_t st, Raw r) const { return value(st, r); }
  // false: the table would pass half full (the caller aborts the query: arena overflow)
  __device__ bool store(uint32_t st, Raw r, float est) {
    uint32_t h = home(st);
    if (r.x != st && r.x != BP_EMPTY) {
      for (h++;; h++) {
        r = __ldcg(tab + (h & mask));
        if (r.x == st || r.x == BP_EMPTY) break;
      }
    }
    if (r.x == BP_EMPTY && ++count > limit) return false;
    tab[h & mask] = make_uint2(st, __float_as_uint(est));
    return true;
  }
};

// * none (layout 3) — forest meshes (the polygon graph incl. off-mesh links has no cycle; exact
//   union-find test at upload).  A state's successors exclude the edge it was entered through
//   (pathfinding.rs:165-166), so on a forest a search enters every polygon once and generates every
//   state once: `try_add_node` (astar.rs:139-160) always finds the entry absent (+inf) and the stale
//   test (astar.rs:172-176) never fires.  Nothing is stored, probed or restored; the finished
//   corridor is staged in place at the top of the node list (see stage_index()).
template <>
struct BestStore<3> {
  typedef uint32_t Raw;
  typedef uint32_t RootRaw;
  __device__ __forceinline__ void init(uint8_t*, const AStarArena&) {}
  __device__ __forceinline__ Raw probe(uint32_t) const { return 0u; }
  __device__ __forceinline__ Raw probe_edge(uint32_t) const { return 0u; }
  __device__ __forceinline__ RootRaw probe_root(uint32_t) const { return 0u; }
  __device__ __forceinline__ float root_value(uint32_t, RootRaw) const { return __builtin_huge_valf(); }
  __device__ __forceinline__ float value(uint32_t, Raw) const { return __builtin_huge_valf(); }
  __device__ __forceinline__ float value_edge(uint32_t, Raw) const { return __builtin_huge_valf(); }
  __device__ __forceinline__ float value_fast(uint32_t, Raw, bool& need_table) const {
    need_table = false;
    return __builtin_huge_valf();
  }
  __device__ __forceinline__ bool store(uint32_t, Raw, float) { return true; }
};

__device__ __forceinline__ uint32_t ovf_used(const BestStore<0>&) { return 0u; }
__device__ __forceinline__ uint32_t ovf_used(const BestStore<1>& b) { return b.ovf_count; }
__device__ __forceinline__ uint32_t ovf_used(const BestStore<2>&) { return 0u; }
__device__ __forceinline__ uint32_t ovf_used(const BestStore<3>&) { return 0u; }
__device__ __forceinline__ void ovf_clear(BestStore<0>&) {}
__device__ __forceinline__ void ovf_clear(BestStore<1>& b) { b.ovf_count = 0; }
__device__ __forceinline__ void ovf_clear(BestStore<2>& b) { b.count = 0; }
__device__ __forceinline__ void ovf_clear(BestStore<3>&) {}

// Restores a finished query's best_estimates region to "absent" (all 32 lanes of a warp cooperate).
// nn nodes were created (nd = the slot's node list); n_staged entries at the start of the region
// were overwritten by the staged corridor.
template <int LAYOUT>
__device__ __forceinline__ void restore_best(const AStarArena& arena, uint8_t* base, const uint2* nd, uint32_t nn,
                                             uint32_t n_staged, bool ovf_used, uint32_t lane) {
  if (LAYOUT == 3) return;
  if (LAYOUT == 2) {
    // hashed: entries cannot be removed one by one (probe chains): clear the whole table
    uint4* t4 = reinterpret_cast<uint4*>(base);
    const uint4 e4 = make_uint4(BP_EMPTY, 0u, BP_EMPTY, 0u);
    for (uint32_t k = lane; k < arena.hash_cap / 2; k += 32) t4[k] = e4;
  } else if (LAYOUT == 0) {
    float* b = reinterpret_cast<float*>(base);
    if ((size_t)nn * 16 >= arena.n_states) {
      float4* b4 = reinterpret_cast<float4*>(b);
      const uint32_t n4 = (arena.n_states + 3) >> 2;
      const float4 inf4 = make_float4(F_INF, F_INF, F_INF, F_INF);
      for (uint32_t k = lane; k < n4; k += 32) b4[k] = inf4;
    } else {
      for (uint32_t k = lane; k < nn; k += 32) b[__ldcg(nd + k).x] = F_INF;
      uint2* st = reinterpret_cast<uint2*>(base);
      const uint2 e2 = make_uint2(__float_as_uint(F_INF), __float_as_uint(F_INF));
      for (uint32_t k = lane; k < n_staged; k += 32) st[k] = e2;
    }
  } else {
    uint2* bp = reinterpret_cast<uint2*>(base);
    float* sp = reinterpret_cast<float*>(base + arena.special_offset);
    const uint32_t n_polys = arena.n_ids >> arena.kshift;
    if ((size_t)nn * 8 >= n_polys) {
      uint4* b4 = reinterpret_cast<uint4*>(base);
      const uint32_t n4 = (n_polys + 1) >> 1;
      const uint4 e4 = make_uint4(__float_as_uint(F_INF), BP_EMPTY, __float_as_uint(F_INF), BP_EMPTY);
      for (uint32_t k = lane; k < n4; k += 32) b4[k] = e4;
      for (uint32_t k = lane; k < arena.n_states - arena.n_ids; k += 32) sp[k] = F_INF;
    } else {
      for (uint32_t k = lane; k < nn; k += 32) {
        const uint32_t st = __ldcg(nd + k).x;
        if (st < arena.n_ids)
          bp[st >> arena.kshift] = make_uint2(__float_as_uint(F_INF), BP_EMPTY);
        else
          sp[st - arena.n_ids] = F_INF;
      }
      const uint2 e2 = make_uint2(__float_as_uint(F_INF), BP_EMPTY);
      for (uint32_t k = lane; k < n_staged; k += 32) bp[k] = e2;
    }
    if (ovf_used) {
      uint4* o4 = reinterpret_cast<uint4*>(base + arena.ovf_offset);
      const uint4 e4 = make_uint4(BP_EMPTY, 0u, BP_EMPTY, 0u);
      for (uint32_t k = lane; k < arena.ovf_cap / 2; k += 32) o4[k] = e4;
    }
  }
}

// Follows the back-pointers of ONE finished search from its End node to the Start node (astar.rs:86-105) with
// all 32 lanes of the warp: the chain is read through a 128-node window held in registers (four nodes per lane,
// moved by shuffles); a dependent memory round trip is paid only when the chain leaves the window.  Calls
// `stage(k, {state, next state})` on lane 0 for the k-th corridor node, goal side first (the End node itself is
// not part of the corridor), and returns the number of corridor nodes.  `nodes` = the slot's node list
// {state, prev}, allocated in 256-byte units (node_cap_window = node_cap rounded up to 32).
template <class StageFn>
__device__ __forceinline__ uint32_t warp_walk(const uint2* nodes, uint32_t node_cap_window, uint32_t goal_index,
                                              uint32_t st_end, uint32_t lane, StageFn stage) {
  uint32_t w_pos = 0, w_node = goal_index, w_next = st_end;
  bool done = false;
  while (!done) {
    const uint32_t base = w_node & ~127u;
    uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;
    const uint32_t i0 = base + 4u * lane;
    if (i0 < node_cap_window) {
      const uint4* sec = reinterpret_cast<const uint4*>(nodes + i0);
      lo = __ldcg(sec);
      hi = __ldcg(sec + 1);
    }
    for (;;) {
      const uint32_t idx = w_node - base, wi = idx & 3u;
      const uint32_t mx = wi == 0 ? lo.x : (wi == 1 ? lo.z : (wi == 2 ? hi.x : hi.z));
      const uint32_t my = wi == 0 ? lo.y : (wi == 1 ? lo.w : (wi == 2 ? hi.y : hi.w));
      const uint32_t rx = __shfl_sync(FULL, mx, idx >> 2), ry = __shfl_sync(FULL, my, idx >> 2);
      if (rx != st_end) {
        if (lane == 0) stage(w_pos, make_uint2(rx, w_next));
        w_pos++;
      }
      w_next = rx;
      if (ry == LMB_NONE) {
        done = true;
        break;
      }
      w_node = ry;
      if ((ry & ~127u) != base) break;
    }
  }
  return w_pos;
}

enum : uint32_t { PH_IDLE = 0, PH_SEARCH = 1, PH_WALK = 2, PH_FINISH = 3, PH_EXIT = 4 };


}  // namespace astar_detail
}  // namespace lmb
