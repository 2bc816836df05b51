// Kernel 2, low-latency form — batched A*, one query per WARP.
//
// Same search, same results as k_astar.cu (pathfinding.rs:19-382, astar.rs:126-206; corridors, portal
// codes and `explored_nodes` bit-exact), same arena slots, same outputs; what differs is who does the
// work.  k_astar.cu gives a query to one thread and hides latency with thousands of queries in flight:
// the best shape when a batch is large, but a single search then runs ~700 dependent warp instructions
// per expansion (2-2.5 us per explored node) and a small batch — the few hundred re-plans of a steady
// frame — takes as long as its longest search.  Here the 32 lanes of a warp share ONE search, and the
// serial chain of an expansion is cut to what really is serial:
//
//   * expansion: lane j owns edge j of the popped polygon — its record (32 B), its row of the distance
//     table, its `best_estimates` probe, its cost / heuristic (two square roots) and its accept test
//     all run side by side; a ballot gives the accepted set in ascending edge order (the reference's
//     successor order), node-list appends and `best` stores are one parallel step;
//   * open list: the exact `BinaryHeap<Reverse<NodeRef>>` emulation of k_astar_common.cuh, but
//       - `sift_down_to_bottom` picks the greater child independently of the hole, so its path is a
//         function of the heap alone: 31 lanes evaluate the winner bits of a five-level subtree at
//         once (two 16-byte shared-memory reads each) and a ballot publishes them; a node lies on the
//         path when every ancestor's bit points its way, which each lane checks against two precomputed
//         masks, and a second ballot IS the path: five levels per round, no walk;
//       - the moves along the path are one parallel step (lane k moves the entry of level k), and the
//         `sift_up` that follows is ONE ballot over the path entries (stop at the first ancestor the
//         hole element is <= to, scanning upward: the highest set bit);
//       - `push` = one parallel read of the ancestors, one ballot (lowest set bit), one parallel write;
//     all with the reference comparator (`rev_le`, including the cost-vs-estimate quirk of astar.rs:71);
//   * the loads of an expansion (records, distance row, best[root]) are issued before the pop and
//     consumed after it, like in k_astar.cu;
//   * walk: the back-pointer chain is followed through a 128-node window held in registers (four
//     nodes per lane, moved by shuffles); a dependent memory round trip is paid only when the chain
//     leaves the window;
//   * two warps per query (PAIR, the default): `pop` needs nothing from the expansion and the expansion
//     only the root's identity, so a HEAP warp owns the open list — it publishes the root, pops it, and
//     pushes the batch of successors it is handed — while the SEARCH warp expands the root; a mailbox in
//     shared memory and one named barrier (bar.sync id, 64) between them.  The search warp spends the push
//     phase guessing the next root (the batch entry that passes every ancestor, else the root the pop
//     left) and fetching its records.
//
// A query unit (warp or pair) runs one query at a time and pulls the next from the batch's queue; 1, 4, 8
// or 16 units per SM with ~13 800 / 3 400 / 1 500 / 700 open-list entries per query in shared memory
// (in-place spill beyond that).
#include <algorithm>

#include "k_astar_common.cuh"

namespace lmb {

using namespace astar_detail;

namespace {

// rev_le on packed entries {cost, estimate, index, state}
__device__ __forceinline__ bool rev_le4(const uint4& x, const uint4& y) {
  const float xc = __uint_as_float(x.x), xe = __uint_as_float(x.y), ye = __uint_as_float(y.y),
              yc = __uint_as_float(y.x);
  (void)xc;
  // Reverse::le(x, y) = NodeRef::le(y, x): y.estimate < x.estimate, or equal estimates and x.estimate <= y.cost
  return (ye < xe) | ((ye == xe) & (xe <= yc));
}

template <uint32_t HS>
struct WarpHeap {
  uint4* sm;      // this warp's entries [0, HS)
  uint4* gm;      // spill: entry i (>= HS) at gm[i - HS + 1]
  uint32_t len;
  uint32_t lane;
  uint32_t lv_m, lv_o;  // level / offset of this lane's node inside a 31-node (five-level) subtree
  uint32_t anc_one, anc_zero;

  __device__ __forceinline__ void init(uint4* sm_, uint4* gm_, uint32_t lane_) {
    sm = sm_;
    gm = gm_;
    len = 0;
    lane = lane_;
    lv_m = 31u - __clz(lane_ + 1u);
    lv_o = lane_ + 1u - (1u << lv_m);
    // ancestors of this lane's node inside the subtree, split by the winner bit that leads towards the node
    anc_one = anc_zero = 0;
    for (uint32_t j = 0; j < lv_m && lane_ < 31u; j++) {
      const uint32_t a = (1u << j) - 1u + (lv_o >> (lv_m - j));
      if ((lv_o >> (lv_m - j - 1u)) & 1u)
        anc_one |= 1u << a;
      else
        anc_zero |= 1u << a;
    }
  }
  // SP = false: every index the operation can reach is < HS (the caller checks the length once, uniformly), so
  // an access is one shared-memory instruction without a branch; SP = true adds the in-place spill.
  template <bool SP>
  __device__ __forceinline__ uint4 ld(uint32_t i) const {
    if (!SP || i < HS) return sm[i];
    return __ldcg(gm + (i - HS + 1));
  }
  template <bool SP>
  __device__ __forceinline__ void st(uint32_t i, const uint4& v) const {
    if (!SP || i < HS)
      sm[i] = v;
    else
      __stcg(gm + (i - HS + 1), v);
  }

  // BinaryHeap::push = sift_up(0, old_len).  Lanes that have nothing to do read entry 0 (a valid address)
  // and are masked out of the ballot: no divergent branch around the loads.
  template <bool SP>
  __device__ __forceinline__ void push_t(const uint4& e) {
    const uint32_t p1 = len + 1u;  // 1-based index of the new position; its k-th ancestor is p1 >> k
    len = p1;
    const uint32_t depth = 31u - __clz(p1);
    const bool has = lane < depth;
    const uint4 anc = ld<SP>(has ? (p1 >> (lane + 1u)) - 1u : 0u);
    const uint32_t mask = __ballot_sync(FULL, has & rev_le4(e, anc));
    const uint32_t rise = mask ? (uint32_t)__ffs(mask) - 1u : depth;  // levels the new entry moves up
    if (lane < rise) st<SP>((p1 >> lane) - 1u, anc);
    if (lane == 0) st<SP>((p1 >> rise) - 1u, e);
    __syncwarp();
  }
  __device__ __forceinline__ void push(const uint4& e) {
    if (len < HS)
      push_t<false>(e);
    else
      push_t<true>(e);
  }

  // BinaryHeap::pop = swap_remove(0) + sift_down_to_bottom(0) + sift_up
  template <bool SP>
  __device__ __forceinline__ uint4 pop_t() {
    len--;
    const uint4 last = ld<SP>(len);
    if (len == 0) return last;
    const uint4 top = sm[0];
    const uint32_t end = len;
    uint32_t pos = 0, d = 0, my_pos = 0;
    while (2u * pos + 2u < end) {  // `pos` has two children
      // winner bits of the five levels below `pos`: lane l < 31 owns the node with local heap index l
      const uint32_t g = ((pos + 1u) << lv_m) + lv_o - 1u;
      const bool two = (lane < 31u) & (2u * g + 2u < end);
      const uint32_t ia = two ? 2u * g + 1u : 0u;
      const uint4 a = ld<SP>(ia), b = ld<SP>(ia + 1u);
      const bool right = two & rev_le4(a, b);  // child += (a <= b): the right child wins
      const uint32_t bits = __ballot_sync(FULL, right);
      // A node lies on the path when every ancestor's bit points its way (a node with two children has ancestors
      // with two children: the heap is a complete tree), so each lane decides for itself and the path is read off
      // a second ballot: one lane per level, levels 0 .. nl-1.
      const bool onp = two & ((bits & anc_one) == anc_one) & ((bits & anc_zero) == 0u);
      const uint32_t pm = __ballot_sync(FULL, onp);
      const uint32_t nl = __popc(pm);
      const uint32_t c = 2u * g + 1u + (right ? 1u : 0u);  // the child this node sends the hole to
      // depth d + m + 1 of the path = the choice of the on-path node of level m; lane (d + m + 1) keeps it
      const uint32_t m = lane - d - 1u, mc = m < 4u ? m : 4u;
      const uint32_t level = ((1u << (1u << mc)) - 1u) << ((1u << mc) - 1u);
      const uint32_t src = (uint32_t)__ffs(pm & level) - 1u;
      const uint32_t v = __shfl_sync(FULL, c, src & 31u);
      if (m < nl) my_pos = v;
      pos = __shfl_sync(FULL, c, 31u - __clz(pm));
      d += nl;
    }
    if (2u * pos + 2u == end) {  // a single child at the bottom
      pos = 2u * pos + 1u;
      d++;
      if (lane == d) my_pos = pos;
    }
    // Entry of level k moves to level k-1 (lane k); the hole element (the old last entry) goes to the
    // bottom and sifts up: it stops below the first entry, scanning upward, that it is <= to.
    const bool on_path = (lane >= 1u) & (lane <= d);
    const uint4 mine = ld<SP>(on_path ? my_pos : 0u);
    const uint32_t mask = __ballot_sync(FULL, on_path & rev_le4(last, mine));
    const uint32_t j = mask ? 31u - __clz(mask) : 0u;
    const uint32_t pos_j = __shfl_sync(FULL, my_pos, j);
    if (lane >= 1u && lane <= j) st<SP>((my_pos - 1u) >> 1, mine);
    if (lane == 0) st<SP>(pos_j, last);
    __syncwarp();
    return top;
  }
  __device__ __forceinline__ uint4 pop() {
    if (len <= HS) return pop_t<false>();
    return pop_t<true>();
  }
};

// ---- asynchronous staging of an expansion's first loads ------------------------------------------------
// The records of the popped polygon, its row of the distance table and best[root] are wanted AFTER the pop,
// and the pop's ~300 cycles of shared-memory work are there to hide their L2 round trip.  Loaded into
// registers they do not overlap it: ptxas gives the pop's first LDS the scoreboard of the outstanding LDGs,
// and the first instruction of the pop then waits the whole round trip (ncu: 25 % of all stall samples on
// that one instruction).  cp.async has no destination register and no scoreboard: global -> shared memory,
// completion by cp.async.wait_group behind the pop.
struct __align__(16) LaneStage {  // per lane
  uint4 a;        // EdgeRec {mx, my, mz, twin}
  uint32_t info;  // EdgeRec::info
  float dist;     // edge_dist[s][lane]
  uint32_t pad[2];
};
struct __align__(16) WarpStage {  // per warp
  LaneStage lane[32];
  uint4 root;  // best[root] probe: layout-specific words
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// best[root] probe through the staging slot: issue (one lane) and read back (all lanes, after the wait)
template <int LAYOUT>
__device__ __forceinline__ void stage_root_probe(const BestStore<LAYOUT>&, uint32_t, uint4*) {}
template <>
__device__ __forceinline__ void stage_root_probe<0>(const BestStore<0>& bs, uint32_t st, uint4* slot) {
  cp_async_4(slot, bs.b + st);
}
template <>
__device__ __forceinline__ void stage_root_probe<1>(const BestStore<1>& bs, uint32_t st, uint4* slot) {
  if (st < bs.n_ids)
    cp_async_8(slot, bs.bp + (st >> bs.kshift));
  else
    cp_async_4(reinterpret_cast<uint32_t*>(slot) + 2, bs.sp + (st - bs.n_ids));
}
template <>
__device__ __forceinline__ void stage_root_probe<2>(const BestStore<2>& bs, uint32_t st, uint4* slot) {
  cp_async_8(slot, bs.tab + (bs.home(st) & bs.mask));
}
template <int LAYOUT>
__device__ __forceinline__ float staged_root_value(const BestStore<LAYOUT>&, uint32_t, const uint4&) {
  return __builtin_huge_valf();
}
template <>
__device__ __forceinline__ float staged_root_value<0>(const BestStore<0>&, uint32_t, const uint4& v) {
  return __uint_as_float(v.x);
}
template <>
__device__ __forceinline__ float staged_root_value<1>(const BestStore<1>& bs, uint32_t st, const uint4& v) {
  if (st >= bs.n_ids) return __uint_as_float(v.z);
  return bs.value_edge(st, make_uint2(v.x, v.y));
}
template <>
__device__ __forceinline__ float staged_root_value<2>(const BestStore<2>& bs, uint32_t st, const uint4& v) {
  return bs.value(st, make_uint2(v.x, v.y));
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// ---- the open list as the search sees it ----------------------------------------------------------------
// PAIR = false: the warp that runs the search also owns the heap (WarpHeap).
// PAIR = true: a second warp — the heap warp — owns it, and the two take turns through a mailbox in shared
// memory and one named barrier.  An expansion's chain pop -> records -> costs -> pushes is serial for one
// warp (~340 dependent-ish instructions per explored node, issue slots 15 % busy); but the pop needs nothing
// from the expansion and the expansion needs only the root's identity, so with two warps they run side by
// side: the heap warp publishes the root and pops it while the search warp expands it; the search warp hands
// the accepted successors over as one batch and the heap warp pushes them in order.  Same heap operations in
// the same order as the single-warp kernel, hence the same corridors.
constexpr uint32_t PAIR_BATCH = 48;
enum : uint32_t { BOX_MORE = 1u, BOX_STOP = 2u, BOX_EXIT = 4u };
struct __align__(16) PairBox {
  uint4 e[PAIR_BATCH];  // entries to push, in order
  uint4 root;           // the root after the pushes (the entry the heap warp then pops)
  uint32_t n, flags;    // batch size; MORE: the expansion goes on, no pop yet; STOP: the query is over; EXIT
  uint32_t len, pad;    // open-list length after the pushes
  uint32_t next_len;    // open-list length before the pushes (after the pop) ...
  uint32_t pad2[3];
  uint4 next;           // ... and its root then: what a pushed entry has to beat to be the next root
};
// The search warp of a pair has no pop to hide the L2 round trip of the root's records behind.  It is idle while
// the heap warp pushes its batch, and it can tell what the root will be then: the entry of the batch that passes
// every ancestor (sift_up stops at the first ancestor the entry is <= to) or else the root the pop left behind.
// It fetches that state's records in the idle window (cp.async into a staging slot, like the single-warp
// kernel's staging), so they are there or nearly there when the root is published.  Records are immutable; a
// wrong guess, a special state or CSR ids are a miss: fetch and wait as before.
constexpr uint32_t REC_CACHE = 1;
struct __align__(16) RecSlot {
  LaneStage lane[32];
};
__device__ __forceinline__ void pair_sync(uint32_t id) {
  __syncwarp();
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <uint32_t HS, bool PAIR>
struct OpenList;

template <uint32_t HS>
struct OpenList<HS, false> {
  WarpHeap<HS> heap;
  __device__ __forceinline__ void init(uint4* sm, uint4* gm, uint32_t lane, PairBox*, uint32_t) { heap.init(sm, gm, lane); }
  __device__ __forceinline__ void reset() { heap.len = 0; }
  __device__ __forceinline__ uint32_t len() const { return heap.len; }
  template <class F>
  __device__ __forceinline__ bool peek(uint32_t& s, F&&) {
    if (heap.len == 0) return false;
    s = heap.sm[0].w;
    return true;
  }
  __device__ __forceinline__ uint4 take() { return heap.pop(); }
  __device__ __forceinline__ void push(const uint4& e) { heap.push(e); }
  // the accepted lanes' entries in ascending lane order; entry k gets node index base + k
  __device__ __forceinline__ void push_lanes(bool, uint32_t amask, uint32_t, float nc, float es, uint32_t state,
                                             uint32_t base) {
    uint32_t m = amask, k = 0;
    while (m) {
      const uint32_t j = (uint32_t)__ffs(m) - 1u;
      m &= m - 1u;
      const uint32_t s2 = __shfl_sync(FULL, state, j);
      const float e2 = __shfl_sync(FULL, es, j), c2 = __shfl_sync(FULL, nc, j);
      heap.push(make_uint4(__float_as_uint(c2), __float_as_uint(e2), base + k, s2));
      k++;
    }
  }
  __device__ __forceinline__ void stop() {}
  __device__ __forceinline__ void exit() {}
};

template <uint32_t HS>
struct OpenList<HS, true> {  // the search warp's end of the mailbox
  PairBox* box;
  uint32_t id, lane, n, len_x;
  uint4 root;
  __device__ __forceinline__ void init(uint4*, uint4*, uint32_t lane_, PairBox* box_, uint32_t id_) {
    box = box_, id = id_, lane = lane_, n = 0, len_x = 0;
  }
  __device__ __forceinline__ void reset() { n = 0, len_x = 0; }
  __device__ __forceinline__ uint32_t len() const { return len_x + n; }
  __device__ __forceinline__ void flush(uint32_t flags) {
    __syncwarp();
    if (lane == 0) box->n = n, box->flags = flags;
    pair_sync(id);  // the heap warp takes the batch ...
    pair_sync(id);  // ... has pushed it, and (unless MORE) has published the root it is popping now
    len_x += n;
    n = 0;
  }
  // `idle(box, n)` runs while the heap warp pushes the n entries just handed over
  template <class F>
  __device__ __forceinline__ bool peek(uint32_t& s, F&& idle) {
    __syncwarp();
    if (lane == 0) box->n = n, box->flags = 0;
    pair_sync(id);
    idle(box, n);
    pair_sync(id);
    n = 0;
    root = box->root;
    const uint32_t l = box->len;
    len_x = l ? l - 1u : 0u;
    s = root.w;
    return l != 0;
  }
  __device__ __forceinline__ uint4 take() { return root; }
  __device__ __forceinline__ void push(const uint4& e) {
    if (n == PAIR_BATCH) flush(BOX_MORE);
    if (lane == 0) box->e[n] = e;
    n++;
  }
  __device__ __forceinline__ void push_lanes(bool acc, uint32_t amask, uint32_t rank, float nc, float es, uint32_t state,
                                             uint32_t base) {
    const uint32_t cnt = __popc(amask);
    if (n + cnt > PAIR_BATCH) flush(BOX_MORE);
    if (acc) box->e[n + rank] = make_uint4(__float_as_uint(nc), __float_as_uint(es), base + rank, state);
    n += cnt;
  }
  __device__ __forceinline__ void signal(uint32_t flags) {
    __syncwarp();
    if (lane == 0) box->n = 0, box->flags = flags;
    pair_sync(id);
    // STOP: a second barrier, so that the header of the next query's first batch cannot be written before the heap
    // warp has read this one (every other header is written after the second barrier of its predecessor)
    if (flags & BOX_STOP) pair_sync(id);
    n = 0, len_x = 0;
  }
  __device__ __forceinline__ void stop() { signal(BOX_STOP); }
  __device__ __forceinline__ void exit() { signal(BOX_EXIT); }
};

// the records of padded-id state `st` (its polygon's records and its row of the distance table) -> a staging
// slot; lanes beyond the polygon's records copy its last one again (same line: no traffic) and drop out at the
// `e < cnt` test of the expansion: no divergent branch around the copies
struct RecSource {
  const EdgeRec* recs;
  const float* edge_dist;
  uint32_t kshift, n_ids;
};
__device__ __forceinline__ void issue_records(LaneStage* dst, const RecSource& src, uint32_t st, uint32_t lane) {
  const uint32_t kmask = (1u << src.kshift) - 1u;
  const uint32_t le = lane < kmask + 1u ? lane : kmask;
  const EdgeRec* rp = src.recs + (st & ~kmask) + le;
  cp_async_16(&dst[lane].a, rp);
  cp_async_4(&dst[lane].info, &rp->info);
  cp_async_4(&dst[lane].dist, src.edge_dist + ((size_t)st << src.kshift) + le);
}

// The heap warp of a pair: pushes the batches it is handed, publishes the root, pops it, publishes what the pop
// left at the root.
template <uint32_t HS>
__device__ __forceinline__ void heap_warp_loop(uint4* sm, uint4* gm, PairBox* box, uint32_t id, uint32_t lane) {
  WarpHeap<HS> heap;
  heap.init(sm, gm, lane);
  if (lane == 0) box->next_len = 0;
  for (;;) {
    pair_sync(id);
    const uint32_t n = box->n, flags = box->flags;
    if (flags & BOX_EXIT) break;
    if (flags & BOX_STOP) {
      heap.len = 0;
      if (lane == 0) box->next_len = 0;
      pair_sync(id);  // (see OpenList::signal)
      continue;
    }
    for (uint32_t k = 0; k < n; k++) heap.push(box->e[k]);
    const bool more = flags & BOX_MORE;
    if (!more && lane == 0) box->root = heap.sm[0], box->len = heap.len;
    pair_sync(id);
    if (!more && heap.len != 0) heap.pop();
    if (lane == 0) box->next = heap.sm[0], box->next_len = heap.len;
  }
}

// key under which two states collide physically in a layout's table (never for the dense array)
template <int LAYOUT>
__device__ __forceinline__ uint32_t collision_key(const BestStore<LAYOUT>&, uint32_t st, uint32_t lane) {
  return 0x80000000u | lane;
}
template <>
__device__ __forceinline__ uint32_t collision_key<1>(const BestStore<1>& bs, uint32_t st, uint32_t) {
  return st >> bs.kshift;  // one entry per polygon
}
template <>
__device__ __forceinline__ uint32_t collision_key<2>(const BestStore<2>& bs, uint32_t st, uint32_t) {
  return bs.home(st) & bs.mask;  // a stale probe of the home slot would overwrite a neighbour's insert
}

template <int LAYOUT>
__device__ __forceinline__ bool needs_chain(const BestStore<LAYOUT>&, uint32_t, const typename BestStore<LAYOUT>::Raw&) {
  return false;
}
template <>
__device__ __forceinline__ bool needs_chain<1>(const BestStore<1>& bs, uint32_t st, const uint2& raw) {
  return raw.y != BP_EMPTY && raw.y != (st & bs.kmask);
}
template <>
__device__ __forceinline__ bool needs_chain<2>(const BestStore<2>&, uint32_t st, const uint2& raw) {
  return raw.x != BP_EMPTY && raw.x != st;
}

}  // namespace

template <uint32_t HS, uint32_t WPB, uint32_t BPS, int LAYOUT, bool SIMPLE, bool PAIR>
__global__ void __launch_bounds__(WPB * (PAIR ? 64 : 32), BPS)
k_astar_warp(DevNav nav, AStarArena arena, AStarQueries q, PathPool pool, AStarCounters* counters,
             const uint32_t* __restrict__ type_raw) {
  extern __shared__ uint4 s_heap[];  // [WPB][HS] open lists, [WPB] WarpStage, then (PAIR) [WPB] PairBox, [WPB][REC_CACHE] RecSlot
  __shared__ float s_cost[LMB_MAX_TYPES];
  constexpr uint32_t THREADS = WPB * (PAIR ? 64 : 32);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = PAIR ? threadIdx.x >> 6 : threadIdx.x >> 5;  // the query unit inside the block
  for (uint32_t t = threadIdx.x; t < nav.n_types; t += THREADS) s_cost[t] = nav.type_cost[t];
  __syncthreads();
  const uint32_t slot = blockIdx.x * WPB + warp;
  if (slot >= arena.n_slots) return;

  typedef typename BestStore<LAYOUT>::Raw BestRaw;
  uint8_t* const slot_base = arena.base + (size_t)slot * arena.slot_stride;
  WarpStage* const stage_w = reinterpret_cast<WarpStage*>(s_heap + (size_t)WPB * HS) + warp;
  PairBox* const box = reinterpret_cast<PairBox*>(reinterpret_cast<WarpStage*>(s_heap + (size_t)WPB * HS) + WPB) + warp;
  RecSlot* const cache = reinterpret_cast<RecSlot*>(reinterpret_cast<PairBox*>(reinterpret_cast<WarpStage*>(s_heap + (size_t)WPB * HS) + WPB) + WPB) +
                         (size_t)warp * REC_CACHE;
  if (PAIR && ((threadIdx.x >> 5) & 1u)) {
    heap_warp_loop<HS>(s_heap + (size_t)warp * HS, reinterpret_cast<uint4*>(slot_base + arena.heap_offset), box, 1u + warp,
                       lane);
    return;
  }
  BestStore<LAYOUT> bs;
  bs.init(slot_base, arena);
  uint2* const nodes = reinterpret_cast<uint2*>(slot_base + arena.nodes_offset);  // {state, prev}
  uint2* const stage = reinterpret_cast<uint2*>(slot_base);
  const uint32_t stage_cap =
      LAYOUT == 3 ? 0xffffffffu
                  : (LAYOUT == 1 ? (arena.n_ids >> arena.kshift) : (LAYOUT == 2 ? arena.hash_cap : arena.n_states / 2));
  OpenList<HS, PAIR> heap;
  heap.init(s_heap + (size_t)warp * HS, reinterpret_cast<uint4*>(slot_base + arena.heap_offset), lane, box, 1u + warp);
  LaneStage* const stage_l = stage_w->lane + lane;

  const EdgeRec* __restrict__ recs = nav.edges;
  const uint32_t heap_cap = arena.heap_cap, node_cap = arena.node_cap;
  const uint32_t node_cap_window = (node_cap + 31u) & ~31u;  // the node list is allocated in 256-byte units
  const uint32_t kshift = arena.kshift, kmask = (1u << kshift) - 1u;
  constexpr bool simple = SIMPLE;  // one polygon type, no per-query overrides: every type cost is cost0
  const float cost0 = s_cost[0];
  const uint32_t ST_LINK0 = arena.n_ids;
  const uint32_t ST_START = arena.n_ids + nav.n_links;
  const uint32_t ST_END = ST_START + 1;
  const uint32_t lt = lanemask_lt();
  const RecSource rec_src{recs, nav.edge_dist, kshift, arena.n_ids};

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(&counters->queue_head, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= q.n) break;
    const uint32_t qid = q.list ? q.list[qi] : qi;
    const uint32_t start_node = q.start_node[qid], end_node = q.end_node[qid];
    const V3 start_point{q.start_point[3 * (size_t)qid], q.start_point[3 * (size_t)qid + 1],
                         q.start_point[3 * (size_t)qid + 2]};
    const V3 end_point{q.end_point[3 * (size_t)qid], q.end_point[3 * (size_t)qid + 1], q.end_point[3 * (size_t)qid + 2]};
    const uint32_t owner = q.owner ? q.owner[qid] : qid;
    uint32_t ov_begin = 0, ov_end = 0;
    if (!SIMPLE && q.override_begin) {
      ov_begin = q.override_begin[owner];
      ov_end = q.override_begin[owner + 1];
    }
    // type cost: override -> archipelago -> 1.0 (pathfinding.rs:73-77)
    auto cost_of = [&](uint32_t t) -> float {
      if (simple) return cost0;
      float c = s_cost[t];
      if (ov_begin != ov_end) {
        const uint32_t raw = type_raw[t];
        for (uint32_t k = ov_begin; k < ov_end; k++)
          if (q.override_type[k] == raw) c = q.override_cost[k];
      }
      return c;
    };
    // PermittedAnimationLinks::is_permitted (agent.rs:178-186)
    auto is_kind_permitted = [&](uint32_t kind) -> bool {
      if (!q.owner_flags || (q.owner_flags[owner] & LMB_AGENT_PERMIT_ALL_LINKS)) return true;
      if (q.permitted_begin)
        for (uint32_t m = q.permitted_begin[owner]; m < q.permitted_begin[owner + 1]; m++)
          if (q.permitted_kind[m] == kind) return true;
      return false;
    };
    // cheapest_type_index_cost (pathfinding.rs:300-314)
    float cheapest = 1.0f;
    for (uint32_t t = 0; t < nav.n_types; t++) {
      const float c = cost_of(t);
      if (nav.type_registered[t] && is_finite(c) && c < cheapest) cheapest = c;
    }
    // are_nodes_connected (nav_data.rs:1022-1087); every lane runs the same (rare) BFS on the same data
    bool connected;
    {
      const uint32_t i1 = nav.poly_island[start_node], i2 = nav.poly_island[end_node];
      if (i1 == i2 && nav.poly_region[start_node] == nav.poly_region[end_node]) {
        connected = true;
      } else {
        const uint32_t r1 = nav.node_region_number[start_node], r2 = nav.node_region_number[end_node];
        connected = r1 != LMB_NONE && r2 != LMB_NONE && nav.region_root[r1] == nav.region_root[r2];
        if (!connected && r1 != LMB_NONE && r2 != LMB_NONE && nav.n_possible_links) {
          uint32_t found = 0;
          if (lane == 0) {
            uint32_t* seen = reinterpret_cast<uint32_t*>(nodes);
            uint32_t* queue = seen + nav.n_region_numbers;
            for (uint32_t r = 0; r < nav.n_region_numbers; r++) seen[r] = 0u;
            uint32_t head = 0, tail = 0;
            seen[r1] = 1u;
            queue[tail++] = r1;
            while (head < tail) {
              const uint32_t r = queue[head++];
              if (r == r2) {
                found = 1;
                break;
              }
              for (uint32_t k = 0; k < nav.n_possible_links; k++) {
                if (nav.possible_link_src[k] != r) continue;
                if (!is_kind_permitted(nav.possible_link_kind[k])) continue;
                const uint32_t dst = nav.possible_link_dst[k];
                if (seen[dst]) continue;
                seen[dst] = 1u;
                queue[tail++] = dst;
              }
            }
          }
          connected = __shfl_sync(FULL, found, 0) != 0;
        }
      }
    }

    uint32_t n_nodes = 0, explored = 0, hmax = 0, ovf_total = 0;
    uint32_t status = 1, success = 0, path_off = 0, path_len = 0;
    uint32_t goal_index = LMB_NONE;
    heap.reset();
    ovf_clear(bs);

    // try_add_node of a special state (Start, End, off-mesh link): the same work on every lane
    auto push_special = [&](uint32_t s2, float ncost, float est, uint32_t prev) {
      const BestRaw raw = bs.probe(s2);
      if (bs.value(s2, raw) <= est) return;
      if (n_nodes >= node_cap || heap.len() >= heap_cap) {
        status = heap.len() >= heap_cap ? 7u : 2u;
        return;
      }
      uint32_t ok = 1;
      if (lane == 0) {
        ok = bs.store(s2, raw, est) ? 1u : 0u;
        nodes[n_nodes] = make_uint2(s2, prev);
      }
      __syncwarp();
      if (LAYOUT == 1 || LAYOUT == 2) {
        ok = __shfl_sync(FULL, ok, 0);
        if (!ok) {
          status = LAYOUT == 1 ? 5u : 2u;
          return;
        }
      }
      heap.push(make_uint4(__float_as_uint(ncost), __float_as_uint(est), n_nodes, s2));
      n_nodes++;
      hmax = max(hmax, heap.len());
    };

    if (connected) {
      const float est = fadd(0.0f, fmul(distance(start_point, end_point), cheapest));
      if (!(F_INF <= est)) push_special(ST_START, 0.0f, est, LMB_NONE);
    }

    // =========================================== search ===========================================
    while (status == 1) {
      // ---- peek the root, issue this expansion's loads, then pop ----
      uint32_t s = 0;
      uint32_t x_pred = LMB_NONE;  // (pair) the state expected at the root after the pushes: records -> the slot
      auto idle = [&](const PairBox* bx, uint32_t n_batch) {
        if (!kshift) return;
        uint4 ref = bx->next;
        bool have = bx->next_len != 0;
        uint32_t pred = have ? ref.w : LMB_NONE;
        for (uint32_t k = 0; k < n_batch; k++) {
          const uint4 e = bx->e[k];
          if (!have || !rev_le4(e, ref)) ref = e, have = true, pred = e.w;
        }
        if (pred >= ST_LINK0) return;  // a special state, or an empty open list
        cp_async_wait_all();           // a wrong guess of the previous node may still be landing in the slot
        issue_records(cache[0].lane, rec_src, pred, lane);
        cp_async_commit();
        x_pred = pred;
      };
      if (!heap.peek(s, idle)) break;
      const bool is_edge = s < ST_LINK0;
      const bool preloaded = is_edge && kshift != 0;
      EdgeRec r;
      r.twin = LMB_NONE, r.info = 0, r.poly = 0, r.first_edge = 0, r.aux = 0;
      r.mx = r.my = r.mz = 0.0f;
      EdgeRec self = r;
      float dtab = 0.0f;
      uint32_t first = 0, cnt = 0;
      __syncwarp();  // nobody still reads the previous expansion's staging slots
      const LaneStage* rs = stage_w->lane;  // where the root's records are staged
      if constexpr (PAIR) {
        // records: fetched in the idle window if the guess was right, else fetched now
        bool fetched = false;
        if (LAYOUT != 3) {
          if (lane == 0) stage_root_probe<LAYOUT>(bs, s, &stage_w->root);
          fetched = true;
        }
        if (preloaded) {
          first = s & ~kmask;
          cnt = kmask + 1u;
          if (s == x_pred) {
            rs = cache[0].lane;
            fetched = true;
          } else {
            issue_records(stage_w->lane, rec_src, s, lane);
            fetched = true;
          }
        } else if (is_edge) {
          self = ld_rec(recs + s);
        }
        if (fetched) cp_async_wait_all();
      } else {
        if (LAYOUT != 3 && lane == 0) stage_root_probe<LAYOUT>(bs, s, &stage_w->root);
        if (is_edge) {
          if (kshift) {
            first = s & ~kmask;
            cnt = kmask + 1u;
            issue_records(stage_w->lane, rec_src, s, lane);
          } else {
            self = ld_rec(recs + s);
          }
        }
      }
      const uint4 cur = heap.take();
      const float cur_cost = __uint_as_float(cur.x), cur_est = __uint_as_float(cur.y);
      if (!PAIR) cp_async_wait_all();
      __syncwarp();  // lanes read each other's slots below (lane 0's info, the state's own midpoint, the root probe)
      uint32_t info0 = 0;
      uint4 oa = make_uint4(0, 0, 0, 0);
      if (preloaded) {
        const uint4 a = rs[lane].a;
        r.mx = __uint_as_float(a.x), r.my = __uint_as_float(a.y), r.mz = __uint_as_float(a.z), r.twin = a.w;
        r.info = rs[lane].info;
        dtab = rs[lane].dist;
        info0 = rs[0].info;
        oa = rs[s & kmask].a;  // the state's own midpoint: only the rare branches below (GoToEnd, off-mesh links) use it
      }

      // ---- resolve (node, point, ignored step) — pathfinding.rs:93-143 ----
      uint32_t poly = 0, cur_type = 0, has_links = 0;
      uint32_t ignore_edge = LMB_NONE, ignore_link = LMB_NONE;
      V3 point{0.0f, 0.0f, 0.0f};
      if (is_edge) {
        if (kshift) {
          poly = s >> kshift;
          cur_type = (info0 >> 8) & 0xffu;
          has_links = (info0 >> 24) & 1u;
          point = {__uint_as_float(oa.x), __uint_as_float(oa.y), __uint_as_float(oa.z)};
        } else {
          first = self.first_edge;
          cnt = self.info & 0xffu;
          poly = self.poly;
          cur_type = (self.info >> 8) & 0xffu;
          has_links = (self.info >> 24) & 1u;
          point = {self.mx, self.my, self.mz};
        }
        ignore_edge = s;
      } else if (s != ST_END) {
        if (s == ST_START) {
          poly = start_node;
          point = start_point;
        } else {
          const uint32_t l = s - ST_LINK0;
          poly = nav.link_dst_node[l];
          const float* pp;
          if (nav.link_kind[l] == LMB_LINK_BOUNDARY) {
            pp = nav.link_portal + 6 * (size_t)l;
            ignore_link = nav.link_reverse[l];
          } else {
            pp = nav.link_dst_portal + 6 * (size_t)l;
          }
          point = midpoint(V3{pp[0], pp[1], pp[2]}, V3{pp[3], pp[4], pp[5]});
        }
        if (kshift) {
          first = poly << kshift;
          cnt = kmask + 1u;
        } else {
          first = nav.poly_edge_begin[poly];
          cnt = nav.poly_edge_begin[poly + 1] - first;
        }
        cur_type = nav.poly_type_dense[poly];
        has_links = nav.node_link_begin[poly + 1] > nav.node_link_begin[poly];
      }
      // stale (astar.rs:172-176)
      if (LAYOUT != 3 && staged_root_value<LAYOUT>(bs, s, stage_w->root) < cur_est) continue;
      explored++;
      if (s == ST_END) {  // goal (astar.rs:179-184)
        goal_index = cur.z;
        break;
      }
      const float cur_type_cost = cost_of(cur_type);
      if (poly == end_node) {
        // the single successor GoToEnd (pathfinding.rs:152-155); h(End) = 0
        const float c = fmul(distance(point, end_point), cur_type_cost);
        const float ncost = fadd(cur_cost, c);
        push_special(ST_END, ncost, fadd(ncost, 0.0f), cur.z);
        continue;
      }

      // ---- node connections in ascending edge index, 32 at a time: lane j takes edge e0 + j ----
      for (uint32_t e0 = 0; e0 < cnt && status == 1; e0 += 32) {
        const uint32_t e = e0 + lane;
        if (e0 > 0 || !preloaded) {
          if (e < cnt)
            r = ld_rec(recs + first + e);
          else
            r.twin = LMB_NONE, r.info = 0;
        }
        const bool ok = e < cnt && r.twin != LMB_NONE && first + e != ignore_edge &&
                        is_finite(cost_of((r.info >> 16) & 0xffu));
        BestRaw praw{};
        if (ok) praw = bs.probe_edge(r.twin);
        const V3 mid{r.mx, r.my, r.mz};
        const float dj = preloaded ? dtab : distance(point, mid);
        const float nc = fadd(cur_cost, fmul(dj, cur_type_cost));
        const float es = fadd(nc, fmul(distance(mid, end_point), cheapest));
        const bool acc = ok && !(bs.value_edge(r.twin, praw) <= es);
        const uint32_t amask = __ballot_sync(FULL, acc);
        if (amask == 0) continue;
        const uint32_t n_acc = __popc(amask);
        if (n_nodes + n_acc > node_cap || heap.len() + n_acc > heap_cap) {
          status = heap.len() + n_acc > heap_cap ? 7u : 2u;
          break;
        }
        const uint32_t rank = __popc(amask & lt);
        if (acc) nodes[n_nodes + rank] = make_uint2(r.twin, cur.z);
        // best_estimates inserts: side by side, except where two accepted states share a physical entry
        // (compact: the same polygon; hashed: the same home slot) — the later one then needs a fresh probe
        if (LAYOUT == 1 || LAYOUT == 2) {
          const uint32_t peers = __match_any_sync(FULL, acc ? collision_key<LAYOUT>(bs, r.twin, lane) : (0x80000000u | lane));
          const bool later = acc && (peers & lt) != 0;
          // an insert that has to walk a probe chain (hashed) or goes to the overflow table (compact) can end
          // anywhere: then every insert of this expansion is made in turn, each with a fresh probe
          const bool chained = acc && needs_chain<LAYOUT>(bs, r.twin, praw);
          uint32_t lmask = __any_sync(FULL, chained) ? amask : __ballot_sync(FULL, later);
          bool fail = false;
          if (acc && !((lmask >> lane) & 1u)) fail = !bs.store(r.twin, praw, es);
          while (lmask) {
            const uint32_t j = (uint32_t)__ffs(lmask) - 1u;
            lmask &= lmask - 1u;
            __syncwarp();
            if (lane == j) fail = !bs.store(r.twin, bs.probe(r.twin), es);
          }
          __syncwarp();
          if (LAYOUT == 1) ovf_total = __reduce_add_sync(FULL, ovf_used(bs));
          if (__any_sync(FULL, fail) || (LAYOUT == 1 && ovf_total > arena.ovf_cap / 2u)) {
            status = LAYOUT == 1 ? 5u : 2u;
            break;
          }
        } else if (LAYOUT == 0) {
          if (acc) bs.store(r.twin, praw, es);
        }
        // pushes in ascending edge order
        heap.push_lanes(acc, amask, rank, nc, es, r.twin, n_nodes);
        n_nodes += n_acc;
        hmax = max(hmax, heap.len());
      }
      // ---- off-mesh links in ascending link index (canonical; pathfinding.rs:196-229) ----
      if (has_links && status == 1) {
        for (uint32_t k = nav.node_link_begin[poly]; k < nav.node_link_begin[poly + 1] && status == 1; k++) {
          const uint32_t l = nav.node_links[k];
          if (l == ignore_link) continue;
          if (!is_finite(cost_of(nav.link_dst_type_dense[l]))) continue;
          float link_cost = 0.0f;
          if (nav.link_kind[l] == LMB_LINK_ANIMATION) {
            if (!is_kind_permitted(nav.link_anim_kind[l])) continue;
            link_cost = nav.link_cost[l];
          }
          const float* pp = nav.link_portal + 6 * (size_t)l;
          const V3 pm = midpoint(V3{pp[0], pp[1], pp[2]}, V3{pp[3], pp[4], pp[5]});
          const float c = fadd(fmul(distance(point, pm), cur_type_cost), link_cost);
          const float ncost = fadd(cur_cost, c);
          V3 hp = pm;  // heuristic of an OffMeshLink state: portal (boundary) / destination portal (animation)
          if (nav.link_kind[l] != LMB_LINK_BOUNDARY) {
            const float* dp = nav.link_dst_portal + 6 * (size_t)l;
            hp = midpoint(V3{dp[0], dp[1], dp[2]}, V3{dp[3], dp[4], dp[5]});
          }
          push_special(ST_LINK0 + l, ncost, fadd(ncost, fmul(distance(hp, end_point), cheapest)), cur.z);
        }
      }
    }

    heap.stop();  // (pair) the heap warp drops what is left of the open list and waits for the next query

    // ============================ walk the back-pointers (astar.rs:86-105) ============================
    // staged goal-first as (state, next state) pairs; k_path_fixup turns them into (node, portal)
    uint32_t w_pos = 0;
    if (status == 1 && goal_index != LMB_NONE) {
      __syncwarp();
      w_pos = warp_walk(nodes, node_cap_window, goal_index, ST_END, lane, [&](uint32_t k, uint2 pair) {
        if (LAYOUT == 3)
          nodes[n_nodes - 1u - k] = pair;
        else if (k < stage_cap)
          stage[k] = pair;
      });
      if (w_pos > stage_cap) {
        status = 6;  // cannot happen for a simple path
      } else {
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(pool.cursor, (unsigned long long)w_pos);
        off = __shfl_sync(FULL, off, 0);
        if (off + w_pos > pool.capacity) {
          status = 3;
        } else {
          success = 1;
          path_off = (uint32_t)off;
          path_len = w_pos;
        }
      }
    }

    // ================= finish: corridor -> pool, best_estimates back to "absent", results =================
    __syncwarp();
    if (success) {
      if (LAYOUT == 3) {
        for (uint32_t k = lane; k < path_len; k += 32)
          pool.entries[(size_t)path_off + k] = __ldcg(nodes + (n_nodes - path_len + k));
      } else {
        for (uint32_t k = lane; k < path_len; k += 32)
          pool.entries[(size_t)path_off + (path_len - 1u - k)] = __ldcg(stage + k);
      }
    }
    __syncwarp();
    if (n_nodes) {
      const uint32_t staged = (status != 2 && status != 5) ? (w_pos < stage_cap ? w_pos : stage_cap) : 0u;
      restore_best<LAYOUT>(arena, slot_base, nodes, n_nodes, staged, ovf_total != 0, lane);
    }
    __syncwarp();
    if (lane == 0) {
      q.out_status[qid] = status;
      if (status == 1) {
        q.out_success[qid] = success;
        q.out_explored[qid] = explored;
        q.out_path_off[qid] = path_off;
        q.out_path_len[qid] = path_len;
        if (q.slot) {
          const uint32_t sl = q.slot[qid];
          pool.slot_path_off[sl] = path_off;
          pool.slot_path_len[sl] = path_len;
          pool.slot_cost_hint[sl] = explored;
        }
        atomicAdd(&counters->n_done, 1u);
        atomicAdd(&counters->explored_total, (unsigned long long)explored);
        atomicAdd(&counters->path_nodes_total, (unsigned long long)path_len);
        atomicAdd(&counters->heap_max_total, (unsigned long long)hmax);
      } else if (status == 2) {
        atomicAdd(&counters->n_retry_arena, 1u);
      } else if (status == 7) {
        atomicAdd(&counters->n_retry_heap, 1u);
      } else if (status == 5) {
        atomicAdd(&counters->n_retry_dense, 1u);
      } else if (status == 6) {
        atomicAdd(&counters->n_failed, 1u);
      } else {
        atomicAdd(&counters->n_retry_pool, 1u);
      }
    }
  }
  heap.exit();
  if (PAIR) cp_async_wait_all();
}

template <uint32_t HS, uint32_t WPB, uint32_t BPS, int LAYOUT, bool SIMPLE, bool PAIR>
static void launch_warp_t(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                          AStarCounters* counters, const uint32_t* type_raw, uint32_t sm_count,
                          cudaStream_t stream) {
  static_assert(HS >= AT_HEAP_SM_MIN, "a slot's open-list spill is sized for HS >= AT_HEAP_SM_MIN (arena_layout)");
  static_assert(!PAIR || WPB <= 15, "one named barrier per pair");
  const size_t smem = (size_t)HS * WPB * sizeof(uint4) + WPB * sizeof(WarpStage) +
                      (PAIR ? WPB * (sizeof(PairBox) + REC_CACHE * sizeof(RecSlot)) : 0);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_astar_warp<HS, WPB, BPS, LAYOUT, SIMPLE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    cudaFuncSetAttribute(k_astar_warp<HS, WPB, BPS, LAYOUT, SIMPLE, PAIR>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  uint32_t slots = arena.n_slots < q.n ? arena.n_slots : q.n;
  const uint32_t resident = sm_count * BPS * WPB;
  if (slots > resident) slots = resident;
  AStarArena a = arena;
  a.n_slots = slots;
  k_astar_warp<HS, WPB, BPS, LAYOUT, SIMPLE, PAIR>
      <<<(slots + WPB - 1) / WPB, WPB * (PAIR ? 64 : 32), smem, stream>>>(nav, a, q, pool, counters, type_raw);
}

// Returns the launch kind: 2 = one warp per query, 3 = a pair of warps per query.
uint32_t launch_astar_warp(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                           AStarCounters* counters, const uint32_t* type_raw, bool simple, bool pair, uint32_t sm_count,
                           cudaStream_t stream) {
  // The fewer queries, the more of the 227 KB of shared memory each open list gets (a spilled level costs an L2
  // round trip per pop): one query per SM -> ~13 800 entries on chip; 4 per SM -> ~3 400; 8 -> ~1 500; 16 -> ~700.
  // Every shape exists as one warp per query and as a warp PAIR per query (search warp + heap warp, see OpenList):
  // a lone search is bound by one warp's dependent instruction chain, and up to 16 queries per SM the SM has the
  // issue slots for a second warp each.
  const uint32_t per_sm = (std::min(q.n, arena.n_slots) + sm_count - 1) / sm_count;
  const bool use_pair = pair;
#define LMB_WARP_LAUNCH_P(HS, WPB, BPS, L, P)                                                                      \
  do {                                                                                                             \
    if (simple) launch_warp_t<HS, WPB, BPS, L, true, P>(nav, arena, q, pool, counters, type_raw, sm_count, stream);  \
    else launch_warp_t<HS, WPB, BPS, L, false, P>(nav, arena, q, pool, counters, type_raw, sm_count, stream);        \
  } while (0)
#define LMB_WARP_LAUNCH(L)                                                \
  do {                                                                    \
    if (per_sm <= 1 && use_pair) LMB_WARP_LAUNCH_P(13760, 1, 1, L, true);  \
    else if (per_sm <= 1) LMB_WARP_LAUNCH_P(13824, 1, 1, L, false);        \
    else if (per_sm <= 4 && use_pair) LMB_WARP_LAUNCH_P(3392, 4, 1, L, true); \
    else if (per_sm <= 4) LMB_WARP_LAUNCH_P(3456, 4, 1, L, false);         \
    else if (per_sm <= 8 && use_pair) LMB_WARP_LAUNCH_P(1600, 4, 2, L, true); \
    else if (per_sm <= 8) LMB_WARP_LAUNCH_P(1472, 4, 2, L, false);         \
    else if (use_pair) LMB_WARP_LAUNCH_P(704, 4, 4, L, true);              \
    else LMB_WARP_LAUNCH_P(736, 4, 4, L, false);                           \
  } while (0)
  if (arena.layout == 1) LMB_WARP_LAUNCH(1);
  else if (arena.layout == 2) LMB_WARP_LAUNCH(2);
  else if (arena.layout == 3) LMB_WARP_LAUNCH(3);
  else LMB_WARP_LAUNCH(0);
#undef LMB_WARP_LAUNCH
#undef LMB_WARP_LAUNCH_P
  return use_pair ? 3u : 2u;
}

}  // namespace lmb
