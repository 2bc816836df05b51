// Kernel 2 — batched A*, one query per THREAD, persistent lanes pulling from a queue.
//
// Replaces `pathfinding::find_path` + `ArchipelagoPathProblem` + `astar::find_path`
// (crates/landmass/src/pathfinding.rs:19-382, astar.rs:126-206).
//
// Parity contract: corridors, portal edge indices and `explored_nodes` are bit-exact.
// The open list is therefore NOT a bucketed queue: it is an exact emulation of Rust's
// std `BinaryHeap<Reverse<NodeRef>>` (push = sift_up, pop = sift_down_to_bottom +
// sift_up) with the reference's comparator, including the cost-vs-estimate quirk at
// astar.rs:71.
//
// Why a thread per query: the search is a serial chain of dependent memory round trips
// (pop -> best[s] / edge records -> best[successor] -> push), so the only parallelism is
// ACROSS queries, and the number of queries in flight is what hides the latency.  A lane
// owns a query; the 32 lanes of a warp run 32 independent searches in lock-step through
// the same pop / expand / push code.  Each lane is a small state machine
// (IDLE -> SEARCH -> WALK -> FINISH -> IDLE) so that a lane that found its goal walks its
// back-pointers (one dependent load per iteration, issued before and consumed after the
// other lanes' expansion so it adds no latency) while its neighbours keep searching.
//
// Memory:
//   * open list: the first AT_HEAP_SM entries of every lane's heap live in shared memory,
//     interleaved by thread (entry i of thread t at [i * STRIDE + t], conflict-free for
//     128-bit accesses); deeper entries spill in place to the lane's arena slot.
//   * `best_estimates` (a HashMap in the reference): dense f32 array per slot, +inf between
//     queries; restored by a warp-cooperative memset (or a scatter over the node list when
//     the query touched little).
//   * node list {state, prev}: append-only, 8 B per node, per slot.
//   * state ids: NodeEdge{node, start_edge} = edge-record id (CSR edge id, or
//     `poly << kshift | local edge` when every polygon has <= 8 edges so that the records of
//     a polygon are addressable from the state id alone), OffMeshLink(l) = n_ids + l,
//     Start = n_ids + n_links, End = Start + 1.
//   * the corridor is written into the path pool as raw (state, next state) pairs by the
//     walk; `k_path_fixup` translates them to (node id, portal code) afterwards, fully
//     parallel.
#include "k_astar_common.cuh"


namespace lmb {

using namespace astar_detail;


// QPW: queries per warp (lanes >= QPW idle: fewer searches in lock-step per warp).
// KS: compile-time kshift (2 = every polygon has <= 4 edges: one chunk, no chunk loop), or 0 =
//     use the run-time value (0 or 3).
// LAYOUT: best_estimates layout (k_astar_common.cuh): 0 dense, 1 compact, 2 hashed.
// HS: open-list entries per query in shared memory; BPS: blocks per SM the variant is sized for.
// SIMPLE: one polygon type and no per-query cost overrides in this batch (the host checks), so
//     every type cost is one value: no cost-table or override loads anywhere in the search.  Those
//     loads share hardware scoreboards with the best[] probes, and their (never taken) paths made
//     the first cost test wait for the probes it was meant to overlap (ncu: 15 % of stall samples).
template <uint32_t QPW, uint32_t HS, uint32_t BPS, uint32_t KS, int LAYOUT, bool SIMPLE>
__global__ void __launch_bounds__(AT_THREADS, BPS)
k_astar(DevNav nav, AStarArena arena, AStarQueries q, PathPool pool, AStarCounters* counters,
        const uint32_t* __restrict__ type_raw) {
  constexpr uint32_t QPB = QPW * (AT_THREADS / 32);  // queries per block
  extern __shared__ uint4 s_heap[];  // [HS][QPB] open-list entries
  __shared__ float s_cost[LMB_MAX_TYPES];
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t qib = (tid >> 5) * QPW + (lane < QPW ? lane : 0);  // query index in block
  const uint32_t slot = lane < QPW ? blockIdx.x * QPB + qib : 0xffffffffu;
  for (uint32_t t = tid; t < nav.n_types; t += AT_THREADS) s_cost[t] = nav.type_cost[t];
  __syncthreads();

  const uint32_t slot_c = slot < arena.n_slots ? slot : 0;  // out-of-range lanes never touch memory
  uint8_t* const slot_base = arena.base + (size_t)slot_c * arena.slot_stride;
  constexpr bool COMPACT = LAYOUT == 1;
  typedef typename BestStore<LAYOUT>::Raw BestRaw;
  typedef typename BestStore<LAYOUT>::RootRaw BestRootRaw;
  BestStore<LAYOUT> bs;
  bs.init(slot_base, arena);
  uint2* const nodes = reinterpret_cast<uint2*>(slot_base + arena.nodes_offset);  // {state, prev}
  Heap<QPB, HS> heap;
  heap.sm = s_heap + qib;
  heap.gm = reinterpret_cast<uint4*>(slot_base + arena.heap_offset);
  // the finished search's corridor is staged in the slot's (no longer needed) best_estimates
  // region as raw (state, next state) pairs, goal first, before it is copied to the path pool
  // (layout 3 has no such region: the pairs are staged in place at the top of the node list, the k-th
  // at nodes[n_nodes - 1 - k].  The walk visits strictly decreasing node indices i_0 > i_1 > ..., so
  // i_k <= n_nodes - 1 - k: a staged pair never lands on a node the walk has yet to read.)
  uint2* const stage = reinterpret_cast<uint2*>(slot_base);
  const uint32_t stage_cap =
      LAYOUT == 3 ? 0xffffffffu
                  : (LAYOUT == 1 ? (arena.n_ids >> arena.kshift) : (LAYOUT == 2 ? arena.hash_cap : arena.n_states / 2));
  heap.len = 0;

  const EdgeRec* __restrict__ recs = nav.edges;
  const uint32_t heap_cap = arena.heap_cap;
  const uint32_t node_cap = arena.node_cap;
  const uint32_t kshift = KS ? KS : arena.kshift;
  const uint32_t kmask = (1u << kshift) - 1u;
  // one polygon type and no per-query overrides in this batch: every type cost is this value
  constexpr bool simple_cost = SIMPLE;
  const float cost0 = s_cost[0];
  const float4* __restrict__ edge_dist = reinterpret_cast<const float4*>(nav.edge_dist);
  const uint32_t ST_LINK0 = arena.n_ids;
  const uint32_t ST_START = arena.n_ids + nav.n_links;
  const uint32_t ST_END = ST_START + 1;

  // ---- lane state ----
  uint32_t phase = slot < arena.n_slots ? PH_IDLE : PH_EXIT;
  uint32_t qid = 0, end_node = 0;
  V3 end_point{0.0f, 0.0f, 0.0f};
  float cheapest = 1.0f;
  uint32_t ov_begin = 0, ov_end = 0;  // this query's cost overrides
  uint32_t n_nodes = 0, explored = 0;
  // walk
  uint32_t w_node = 0, w_next = 0, w_pos = 0;
  uint32_t hmax = 0;  // largest open list of this query (sizes the next batch's launch shape)

  // type cost: override -> archipelago -> 1.0 (pathfinding.rs:73-77)
  auto cost_of = [&](uint32_t t) -> float {
    if (simple_cost) return cost0;
    float c = s_cost[t];
    if (ov_begin != ov_end) {
      const uint32_t raw = type_raw[t];
      for (uint32_t k = ov_begin; k < ov_end; k++)
        if (q.override_type[k] == raw) c = q.override_cost[k];
    }
    return c;
  };
  // PermittedAnimationLinks::is_permitted (agent.rs:178-186)
  // (owner and its permit-all flag are looked up here, in the rare link paths, instead of being carried
  // in registers through every iteration: the kernel is at its register limit)
  auto is_kind_permitted = [&](uint32_t kind) -> bool {
    const uint32_t owner = q.owner ? q.owner[qid] : qid;
    if (!q.owner_flags || (q.owner_flags[owner] & LMB_AGENT_PERMIT_ALL_LINKS)) return true;
    if (q.permitted_begin)
      for (uint32_t m = q.permitted_begin[owner]; m < q.permitted_begin[owner + 1]; m++)
        if (q.permitted_kind[m] == kind) return true;
    return false;
  };
  // first edge record / record count of a polygon
  auto poly_span = [&](uint32_t poly, uint32_t& first, uint32_t& cnt) {
    if (kshift) {
      first = poly << kshift;
      cnt = kmask + 1u;
    } else {
      first = nav.poly_edge_begin[poly];
      cnt = nav.poly_edge_begin[poly + 1] - first;
    }
  };

  for (;;) {
    // results of a query that finishes in this iteration (every path into FINISH sets them in the
    // iteration that publishes them, so they are not carried across iterations)
    uint32_t status = 1, success = 0, path_off = 0, path_len = 0;
    // =========================== IDLE: fetch and set up a query ===========================
    if (phase == PH_IDLE) {
      const uint32_t qi = atomicAdd(&counters->queue_head, 1u);
      if (qi >= q.n) {
        phase = PH_EXIT;
      } else {
        qid = q.list ? q.list[qi] : qi;
        const uint32_t start_node = q.start_node[qid];
        end_node = q.end_node[qid];
        const V3 start_point{q.start_point[3 * (size_t)qid], q.start_point[3 * (size_t)qid + 1],
                             q.start_point[3 * (size_t)qid + 2]};
        end_point = {q.end_point[3 * (size_t)qid], q.end_point[3 * (size_t)qid + 1],
                     q.end_point[3 * (size_t)qid + 2]};
        const uint32_t owner = q.owner ? q.owner[qid] : qid;
        ov_begin = ov_end = 0;
        if (!SIMPLE && q.override_begin) {
          ov_begin = q.override_begin[owner];
          ov_end = q.override_begin[owner + 1];
        }
        // cheapest_type_index_cost (pathfinding.rs:300-314)
        cheapest = 1.0f;
        for (uint32_t t = 0; t < nav.n_types; t++) {
          float c = cost_of(t);
          if (nav.type_registered[t] && is_finite(c) && c < cheapest) cheapest = c;
        }
        // are_nodes_connected (nav_data.rs:1022-1087)
        bool connected;
        {
          uint32_t i1 = nav.poly_island[start_node], i2 = nav.poly_island[end_node];
          if (i1 == i2 && nav.poly_region[start_node] == nav.poly_region[end_node]) {
            connected = true;
          } else {
            uint32_t r1 = nav.node_region_number[start_node], r2 = nav.node_region_number[end_node];
            connected = r1 != LMB_NONE && r2 != LMB_NONE && nav.region_root[r1] == nav.region_root[r2];
            if (!connected && r1 != LMB_NONE && r2 != LMB_NONE && nav.n_possible_links) {
              // BFS over region numbers through the permitted PossibleRegionLinks
              // (nav_data.rs:1056-1086).  seen[] / queue[] borrow the (still empty) node list.
              uint32_t* seen = reinterpret_cast<uint32_t*>(nodes);
              uint32_t* queue = seen + nav.n_region_numbers;
              for (uint32_t r = 0; r < nav.n_region_numbers; r++) seen[r] = 0u;
              uint32_t head = 0, tail = 0;
              seen[r1] = 1u;
              queue[tail++] = r1;
              while (head < tail) {
                uint32_t r = queue[head++];
                if (r == r2) {
                  connected = true;
                  break;
                }
                for (uint32_t k = 0; k < nav.n_possible_links; k++) {
                  if (nav.possible_link_src[k] != r) continue;
                  if (!is_kind_permitted(nav.possible_link_kind[k])) continue;
                  uint32_t dst = nav.possible_link_dst[k];
                  if (seen[dst]) continue;
                  seen[dst] = 1u;
                  queue[tail++] = dst;
                }
              }
            }
          }
        }
        explored = 0;
        n_nodes = 0;
        heap.len = 0;
        phase = PH_FINISH;
        if (connected) {
          // try_add_node(Start)
          float est = fadd(0.0f, fmul(distance(start_point, end_point), cheapest));
          if (!(F_INF <= est)) {
            bs.store(ST_START, bs.probe(ST_START), est);
            nodes[0] = make_uint2(ST_START, LMB_NONE);
            n_nodes = 1;
            heap.push({0.0f, est, 0u, ST_START});
            phase = PH_SEARCH;
          }
        }
      }
    }
    if (__all_sync(FULL, phase == PH_EXIT)) break;

    // ============ WALK (part 1): issue this iteration's back-pointer load early ============
    uint2 w_one = make_uint2(0u, 0u);
    if (phase == PH_WALK) w_one = nodes[w_node];

    // =============================== SEARCH: one pop ===============================
    if (phase == PH_SEARCH) {
      if (heap.len == 0) {
        phase = PH_FINISH;  // no path
      } else {
        // ---- peek the root and issue this expansion's first loads BEFORE the pop's sift work:
        //      best[s] (stale test, HBM) and, with padded ids, the polygon's edge records (L2) ----
        const uint32_t s = heap.sm[0].w;
        const BestRootRaw raw_s = bs.probe_root(s);
        const bool is_edge = s < ST_LINK0;
        EdgeRec r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r[j].twin = LMB_NONE, r[j].info = 0, r[j].mx = r[j].my = r[j].mz = 0.0f;
        EdgeRec self;
        self.poly = self.first_edge = self.info = 0;
        self.mx = self.my = self.mz = 0.0f;
        uint32_t first = 0, cnt = 0;
        const bool preloaded = is_edge && kshift;
        float4 dtab = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // distances state midpoint -> edge midpoints
        if (is_edge) {
          // KS == 2: the state's own record is one of the four loaded below, so it is not fetched on its
          // own: a separate load costs a wait in FRONT of the pop (its `info` word was consumed at once, and
          // the register of its unused `twin` word was recycled inside the pop's sift loop — a write that
          // has to wait for the load), which is exactly the latency the early issue is there to hide.
          if (KS != 2) self = ld_rec(recs + s);
          if (kshift) {
            first = s & ~kmask;
            cnt = kmask + 1u;
#pragma unroll
            for (int j = 0; j < 4; j++) r[j] = ld_rec(recs + first + j);
            dtab = __ldg(edge_dist + ((size_t)s << (kshift - 2)));
          }
        }

        const HeapEntry cur = heap.pop();
        // Nothing that consumes the loads above may be scheduled before the pop (the compiler does not
        // know the sift is long and hoists e.g. the has-links test right behind its load, which then
        // waits a full round trip in front of the pop): pass the loaded registers through an opaque
        // statement the pop's shared-memory traffic cannot move across.
        if (KS == 2) {
          asm volatile(""
                       : "+r"(r[0].twin), "+r"(r[1].twin), "+r"(r[2].twin), "+r"(r[3].twin), "+r"(r[0].info),
                         "+r"(r[1].info), "+r"(r[2].info), "+r"(r[3].info)
                       :
                       : "memory");
        }

        // ---- resolve (node, point, ignore step) — pathfinding.rs:93-143 ----
        uint32_t poly = 0, cur_type = 0, has_links = 0;
        uint32_t ignore_edge = LMB_NONE, ignore_link = LMB_NONE;
        V3 point{0.0f, 0.0f, 0.0f};
        if (is_edge) {
          if (KS == 2) {
            // own type / has-links are the same in every record of the polygon; `point` (the state's own
            // midpoint) is only needed by the rare branches below, which pick it out of r[] themselves
            poly = s >> 2;
            cur_type = (r[0].info >> 8) & 0xffu;
            has_links = (r[0].info >> 24) & 1u;
          } else {
            if (!kshift) {
              first = self.first_edge;
              cnt = self.info & 0xffu;
            }
            poly = self.poly;
            cur_type = (self.info >> 8) & 0xffu;
            has_links = (self.info >> 24) & 1u;
            point = {self.mx, self.my, self.mz};
          }
          ignore_edge = s;
        } else if (s != ST_END) {
          if (s == ST_START) {
            poly = q.start_node[qid];
            point = {q.start_point[3 * (size_t)qid], q.start_point[3 * (size_t)qid + 1],
                     q.start_point[3 * (size_t)qid + 2]};
          } else {
            uint32_t l = s - ST_LINK0;
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
          poly_span(poly, first, cnt);
          cur_type = nav.poly_type_dense[poly];
          has_links = nav.node_link_begin[poly + 1] > nav.node_link_begin[poly];
        }
        const float cur_cost = cur.cost;
        const float cur_type_cost = cost_of(cur_type);
        const bool to_end = poly == end_node;  // single successor GoToEnd (pathfinding.rs:152-155)

        // ---- node connections in ascending edge index, four records at a time.  Per chunk:
        //      records (L2) -> `best` probes of every candidate issued together (HBM, speculative:
        //      loads only) -> costs/estimates of all four evaluated unconditionally (independent
        //      chains, no divergence) -> stale test -> accepted successors pushed in edge order ----
        bool overflow = false, live = true, go_dense = false, heap_full = false;
        for (uint32_t e0 = 0; e0 == 0 || (KS != 2 && e0 < cnt && live && !overflow); e0 += 4) {
          if (e0 > 0 || !preloaded) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              if (e0 + j < cnt && s != ST_END)
                r[j] = ld_rec(recs + first + e0 + j);
              else
                r[j].twin = LMB_NONE;
            }
            if (preloaded) dtab = __ldg(edge_dist + ((size_t)s << (kshift - 2)) + (e0 >> 2));
          }
          // type costs of the four neighbours first: whatever these read (cost table, overrides) is
          // consumed before the probes are in flight
          uint32_t passable = 0;
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (is_finite(cost_of((r[j].info >> 16) & 0xffu))) passable |= 1u << j;
          BestRaw praw[4];
          uint32_t cand = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const bool ok = r[j].twin != LMB_NONE && first + e0 + j != ignore_edge && !to_end && ((passable >> j) & 1u);
            praw[j] = ok ? bs.probe_edge(r[j].twin) : BestRaw{};
            cand |= ok ? (1u << j) : 0u;
          }
          // distance(point, edge midpoint): a constant of the mesh for edge states (table built at
          // upload with the same rounding sequence), computed here for Start / link states
          float dj[4];
          if (preloaded) {
            dj[0] = dtab.x, dj[1] = dtab.y, dj[2] = dtab.z, dj[3] = dtab.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) dj[j] = distance(point, V3{r[j].mx, r[j].my, r[j].mz});
          }
          float nc[4], es[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const V3 mid{r[j].mx, r[j].my, r[j].mz};
            const float c = fmul(dj[j], cur_type_cost);
            nc[j] = fadd(cur_cost, c);
            es[j] = fadd(nc[j], fmul(distance(mid, end_point), cheapest));
          }
          if (e0 == 0) {
            // (cur.state == s: naming the popped entry keeps this use behind the pop)
            if (bs.root_value(cur.state, raw_s) < cur.estimate) {  // stale (astar.rs:172-176)
              live = false;
              break;
            }
            explored++;
            if (s == ST_END) {
              // goal: walk the back-pointers (astar.rs:86-105); the End node's record is loaded
              // next iteration
              w_node = cur.index;
              w_next = ST_END;
              w_pos = 0;
              phase = PH_WALK;
              live = false;
              break;
            }
            if (to_end) {
              // h(End) = 0
              if (KS == 2 && is_edge)
                point = {sel4(s & 3u, r[0].mx, r[1].mx, r[2].mx, r[3].mx), sel4(s & 3u, r[0].my, r[1].my, r[2].my, r[3].my),
                         sel4(s & 3u, r[0].mz, r[1].mz, r[2].mz, r[3].mz)};
              float c = fmul(distance(point, end_point), cur_type_cost);
              float ncost = fadd(cur_cost, c);
              float est = fadd(ncost, 0.0f);
              const BestRaw raw_e = bs.probe(ST_END);
              if (!(bs.value(ST_END, raw_e) <= est)) {
                if (n_nodes >= node_cap || heap.len >= heap_cap) {
              heap_full = heap.len >= heap_cap;
                  overflow = true;
                } else {
                  if (!bs.store(ST_END, raw_e, est)) {
                    go_dense = true;
                  } else {
                    nodes[n_nodes] = make_uint2(ST_END, cur.index);
                    heap.push({ncost, est, n_nodes, ST_END});
                    hmax = max(hmax, heap.len);
                    n_nodes++;
                  }
                }
              }
              live = false;  // no other successor
              break;
            }
          }
          // accept = candidate and not (best <= estimate); the common cases of the probe (tracked /
          // absent) are resolved without branches, the overflow table (compact layout) behind one
          uint32_t acc = 0, slow = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            bool need_table = false;
            const float bv = bs.value_fast(r[j].twin, praw[j], need_table);
            acc |= !(bv <= es[j]) ? (1u << j) : 0u;
            slow |= need_table ? (1u << j) : 0u;
          }
          acc &= cand;
          slow &= cand;
          if (slow) {
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (((slow >> j) & 1u) && bs.value_edge(r[j].twin, praw[j]) <= es[j]) acc &= ~(1u << j);
          }
          // The lanes of the warp step through their k-th accepted successor together
          // (ascending j = ascending edge index).
          bool pushed = false;
          while (acc) {
            const uint32_t j = __ffs(acc) - 1;
            acc &= acc - 1;
            if (n_nodes >= node_cap || heap.len >= heap_cap) {
              heap_full = heap.len >= heap_cap;
              overflow = true;
              break;
            }
            const uint32_t s2 = sel4(j, r[0].twin, r[1].twin, r[2].twin, r[3].twin);
            const float est = sel4(j, es[0], es[1], es[2], es[3]);
            const float ncost = sel4(j, nc[0], nc[1], nc[2], nc[3]);
            // The probes were issued together, before this expansion's inserts.  In the compact and hashed
            // layouts two states can share a physical entry (the same polygon; the same home slot or probe
            // chain), so after the first insert a later successor's probe may be out of date: take a fresh one.
            BestRaw raw2 = sel4(j, praw[0], praw[1], praw[2], praw[3]);
            if ((LAYOUT == 1 || LAYOUT == 2) && pushed) raw2 = bs.probe_edge(s2);
            pushed = true;
            if (!bs.store(s2, raw2, est)) {
              go_dense = true;
              break;
            }
            nodes[n_nodes] = make_uint2(s2, cur.index);
            heap.push({ncost, est, n_nodes, s2});
            hmax = max(hmax, heap.len);
            n_nodes++;
          }
          if (go_dense) break;
        }
        // ---- off-mesh links in ascending link index (canonical; pathfinding.rs:196-229) ----
        if (live && has_links && !overflow && !go_dense) {
          if (KS == 2 && is_edge) {  // rare: re-read the own midpoint rather than keep r[] alive across the pushes
            const EdgeRec own = ld_rec(recs + s);
            point = {own.mx, own.my, own.mz};
          }
          for (uint32_t k = nav.node_link_begin[poly]; k < nav.node_link_begin[poly + 1]; k++) {
            uint32_t l = nav.node_links[k];
            if (l == ignore_link) continue;
            if (!is_finite(cost_of(nav.link_dst_type_dense[l]))) continue;
            float link_cost = 0.0f;
            if (nav.link_kind[l] == LMB_LINK_ANIMATION) {
              if (!is_kind_permitted(nav.link_anim_kind[l])) continue;
              link_cost = nav.link_cost[l];
            }
            const float* pp = nav.link_portal + 6 * (size_t)l;
            V3 pm = midpoint(V3{pp[0], pp[1], pp[2]}, V3{pp[3], pp[4], pp[5]});
            float c = fadd(fmul(distance(point, pm), cur_type_cost), link_cost);
            float ncost = fadd(cur_cost, c);
            // heuristic of OffMeshLink state: portal (boundary) / destination portal (animation)
            V3 hp = pm;
            if (nav.link_kind[l] != LMB_LINK_BOUNDARY) {
              const float* dp = nav.link_dst_portal + 6 * (size_t)l;
              hp = midpoint(V3{dp[0], dp[1], dp[2]}, V3{dp[3], dp[4], dp[5]});
            }
            float est = fadd(ncost, fmul(distance(hp, end_point), cheapest));
            uint32_t s2 = ST_LINK0 + l;
            const BestRaw raw_l = bs.probe(s2);
            if (bs.value(s2, raw_l) <= est) continue;
            if (n_nodes >= node_cap || heap.len >= heap_cap) {
              heap_full = heap.len >= heap_cap;
              overflow = true;
              break;
            }
            if (!bs.store(s2, raw_l, est)) {
              go_dense = true;
              break;
            }
            nodes[n_nodes] = make_uint2(s2, cur.index);
            heap.push({ncost, est, n_nodes, s2});
            hmax = max(hmax, heap.len);
            n_nodes++;
          }
        }
        if (overflow || go_dense) {
          // 5: compact layout's overflow table full; 7: open-list spill full; 2: node list (or hashed table) full
          status = (go_dense && COMPACT) ? 5 : ((overflow && heap_full) ? 7 : 2);
          phase = PH_FINISH;
        }
      }
    } else if (phase == PH_WALK) {
      // ============ WALK (part 2): consume the back-pointer loaded above ============
      // (astar.rs:86-105) staged goal-first as (state, next state), one node per warp iteration.
      // (Measured in round 2 on the bench / on config 5 with 10 000 agents: this form 338.5 / 905 ms; round 1's
      // form — the node's whole 32-byte sector, up to four steps per iteration — 337.4 / 950 ms; all 32 lanes
      // walking one finished search at a time through a register window (warp_walk) 386.9 / 888 ms: 17 % fewer
      // instructions, but the warp's other searches stand still during a walk whose every second step is a
      // dependent HBM round trip.)
      if (w_one.x != ST_END) {
        if (LAYOUT == 3)
          nodes[n_nodes - 1 - w_pos] = make_uint2(w_one.x, w_next);
        else if (w_pos < stage_cap)
          stage[w_pos] = make_uint2(w_one.x, w_next);
        w_pos++;
      }
      w_next = w_one.x;
      w_node = w_one.y;
      if (w_one.y == LMB_NONE) {
        // reached the Start node: the corridor has w_pos nodes; allocate it in the pool
        phase = PH_FINISH;
        if (w_pos > stage_cap) {
          status = 6;  // cannot happen for a simple path (corridor longer than the staging area)
        } else {
          const unsigned long long off = atomicAdd(pool.cursor, (unsigned long long)w_pos);
          if (off + w_pos > pool.capacity) {
            status = 3;
          } else {
            success = 1;
            path_off = (uint32_t)off;
            path_len = w_pos;
          }
        }
      }
    }

    // ====== FINISH: restore best_estimates to +inf (warp-cooperative), publish results ======
    uint32_t fin = __ballot_sync(FULL, phase == PH_FINISH);
    while (fin) {
      const int src = __ffs(fin) - 1;
      fin &= fin - 1;
      const uint32_t nn = __shfl_sync(FULL, n_nodes, src);
      if (nn == 0) continue;
      const uint32_t src_slot = __shfl_sync(FULL, slot, src);
      uint8_t* base = arena.base + (size_t)src_slot * arena.slot_stride;
      const uint2* nd = reinterpret_cast<const uint2*>(base + arena.nodes_offset);
      // corridor: staged goal-first in the best_estimates region -> path pool, start-first
      const uint32_t cL = __shfl_sync(FULL, success ? path_len : 0u, src);
      const uint32_t staged = __shfl_sync(FULL, phase == PH_FINISH && status != 2 && status != 5 ? w_pos : 0u, src);
      if (cL) {
        const uint32_t cOff = __shfl_sync(FULL, path_off, src);
        const uint2* st = reinterpret_cast<const uint2*>(base);
        if (LAYOUT == 3) {
          for (uint32_t k = lane; k < cL; k += 32) pool.entries[(size_t)cOff + k] = nd[nn - cL + k];
        } else {
          for (uint32_t k = lane; k < cL; k += 32) pool.entries[(size_t)cOff + (cL - 1 - k)] = st[k];
        }
      } else {
        (void)__shfl_sync(FULL, path_off, src);
      }
      __syncwarp();
      const uint32_t n_staged = staged < stage_cap ? staged : stage_cap;
      if (LAYOUT == 3) {
        // nothing to restore
      } else if (LAYOUT == 2) {
        // hashed: entries cannot be removed one by one (probe chains): clear the whole table
        uint4* t4 = reinterpret_cast<uint4*>(base);
        const uint4 e4 = make_uint4(BP_EMPTY, 0u, BP_EMPTY, 0u);
        for (uint32_t k = lane; k < arena.hash_cap / 2; k += 32) t4[k] = e4;
      } else if (LAYOUT == 0) {
        float* b = reinterpret_cast<float*>(base);
        if ((size_t)nn * 16 >= arena.n_states) {
          // dense restore: coalesced 16-byte stores over the whole array (padded to 16 B)
          float4* b4 = reinterpret_cast<float4*>(b);
          const uint32_t n4 = (arena.n_states + 3) >> 2;
          const float4 inf4 = make_float4(F_INF, F_INF, F_INF, F_INF);
          for (uint32_t k = lane; k < n4; k += 32) b4[k] = inf4;
        } else {
          for (uint32_t k = lane; k < nn; k += 32) b[nd[k].x] = F_INF;
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
            const uint32_t st = nd[k].x;
            if (st < arena.n_ids)
              bp[st >> arena.kshift] = make_uint2(__float_as_uint(F_INF), BP_EMPTY);
            else
              sp[st - arena.n_ids] = F_INF;
          }
          const uint2 e2 = make_uint2(__float_as_uint(F_INF), BP_EMPTY);
          for (uint32_t k = lane; k < n_staged; k += 32) bp[k] = e2;
        }
        if (__shfl_sync(FULL, ovf_used(bs), src)) {
          uint4* o4 = reinterpret_cast<uint4*>(base + arena.ovf_offset);
          const uint4 e4 = make_uint4(BP_EMPTY, 0u, BP_EMPTY, 0u);
          for (uint32_t k = lane; k < arena.ovf_cap / 2; k += 32) o4[k] = e4;
        }
      }
    }
    __syncwarp();
    if (phase == PH_FINISH) {
      q.out_status[qid] = status;
      if (status == 1) {
        q.out_success[qid] = success;
        q.out_explored[qid] = explored;
        q.out_path_off[qid] = path_off;
        q.out_path_len[qid] = path_len;
        if (q.slot) {
          uint32_t sl = q.slot[qid];
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
      n_nodes = 0;
      w_pos = 0;
      hmax = 0;
      ovf_clear(bs);
      phase = PH_IDLE;
    }
  }
}

// Translates the raw (state, next state) pairs the walk left in the pool into
// (node id, portal code) — pathfinding.rs:324-381.  One warp per finished query.
__global__ void k_path_fixup(DevNav nav, AStarArena arena, AStarQueries q, PathPool pool) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ST_LINK0 = arena.n_ids;
  const uint32_t ST_END = arena.n_ids + nav.n_links + 1;
  for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < q.n; w += warps) {
    const uint32_t qid = q.list ? q.list[w] : w;
    if (q.out_status[qid] != 1 || !q.out_success[qid]) continue;
    const uint32_t off = q.out_path_off[qid], L = q.out_path_len[qid];
    const uint32_t start_node = q.start_node[qid];
    for (uint32_t j = lane; j < L; j += 32) {
      const uint2 raw = pool.entries[(size_t)off + j];
      uint32_t node_id;
      if (j == 0)
        node_id = start_node;  // the Start state
      else if (raw.x < ST_LINK0)
        node_id = nav.edges[raw.x].poly;
      else
        node_id = nav.link_dst_node[raw.x - ST_LINK0];
      uint32_t portal;
      if (raw.y == ST_END)
        portal = LMB_NONE;
      else if (raw.y < ST_LINK0)
        portal = nav.edges[raw.y].aux & 0xffu;  // index of the twin edge inside `node_id`
      else
        portal = LMB_PORTAL_LINK | (raw.y - ST_LINK0);
      pool.entries[(size_t)off + j] = make_uint2(node_id, portal);
    }
  }
}

// Longest-processing-time-first order of a batch.  A batch takes at least as long as its longest
// search, and queries are handed to lanes in queue order, so searches expected to be long should
// start first.  The only predictor available is the agent's previous search (a hint: the order
// never changes a result).  Single block: 1024-bucket counting sort on a log scale.
__global__ void __launch_bounds__(1024)
k_order_queries(uint32_t n, const uint32_t* __restrict__ q_slot, const uint32_t* __restrict__ hint,
                uint32_t unknown_hint, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_count[1024];
  __shared__ uint32_t s_scan[1024];
  const uint32_t t = threadIdx.x;
  s_count[t] = 0;
  __syncthreads();
  auto bucket_of = [&](uint32_t i) -> uint32_t {
    uint32_t h = hint[q_slot[i]];
    if (h == 0) h = unknown_hint;
    int b = (int)(log2f((float)h + 1.0f) * 40.0f);
    b = b < 0 ? 0 : (b > 1023 ? 1023 : b);
    return 1023u - (uint32_t)b;  // descending
  };
  for (uint32_t i = t; i < n; i += 1024) atomicAdd(&s_count[bucket_of(i)], 1u);
  __syncthreads();
  // exclusive scan of 1024 counters (Hillis-Steele)
  uint32_t v = s_count[t];
  s_scan[t] = v;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    uint32_t add = t >= off ? s_scan[t - off] : 0u;
    __syncthreads();
    s_scan[t] += add;
    __syncthreads();
  }
  s_count[t] = s_scan[t] - v;  // start offset of bucket t
  __syncthreads();
  for (uint32_t i = t; i < n; i += 1024) order[atomicAdd(&s_count[bucket_of(i)], 1u)] = i;
}

void launch_order_queries(uint32_t n, const uint32_t* q_slot, const uint32_t* slot_cost_hint, uint32_t unknown_hint,
                          uint32_t* order, cudaStream_t stream) {
  if (n) k_order_queries<<<1, 1024, 0, stream>>>(n, q_slot, slot_cost_hint, unknown_hint, order);
}

// distance(midpoint of record s, midpoint of record j of the same polygon) for padded ids:
// out[s * K + j].  Same device functions as the search, so the values are the ones the search
// would compute (pathfinding.rs:178-182).
__global__ void k_build_edge_dist(const EdgeRec* __restrict__ recs, uint32_t n_ids, uint32_t kshift, float* out) {
  const uint32_t K = 1u << kshift;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n_ids * K;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t s = (uint32_t)(i >> kshift), j = (uint32_t)(i & (K - 1));
    const EdgeRec a = recs[s], b = recs[(s & ~(K - 1)) + j];
    out[i] = distance(V3{a.mx, a.my, a.mz}, V3{b.mx, b.my, b.mz});
  }
}

void launch_build_edge_dist(const EdgeRec* recs, uint32_t n_ids, uint32_t kshift, float* out, cudaStream_t stream) {
  if (!kshift || !n_ids) return;
  k_build_edge_dist<<<148 * 4, 256, 0, stream>>>(recs, n_ids, kshift, out);
}

__global__ void k_arena_init(AStarArena arena) {
  // every slot: the best_estimates region to "absent", the overflow table (compact) to empty
  const size_t words = arena.best_bytes / 4;
  const size_t total = (size_t)arena.n_slots * words;
  const uint32_t n_poly_words = arena.layout == 1 ? 2u * (arena.n_ids >> arena.kshift) : 0u;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t slot = i / words, k = i % words;
    uint32_t v = __float_as_uint(F_INF);
    if (arena.layout == 1 && k < n_poly_words && (k & 1)) v = BP_EMPTY;
    if (arena.layout == 2) v = (k & 1) ? 0u : BP_EMPTY;
    reinterpret_cast<uint32_t*>(arena.base + slot * arena.slot_stride)[k] = v;
  }
  if (arena.layout == 1) {
    const size_t ow = (size_t)arena.ovf_cap * 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)arena.n_slots * ow;
         i += (size_t)gridDim.x * blockDim.x) {
      const size_t slot = i / ow, k = i % ow;
      reinterpret_cast<uint32_t*>(arena.base + slot * arena.slot_stride + arena.ovf_offset)[k] = (k & 1) ? 0u : BP_EMPTY;
    }
  }
}

void launch_arena_init(const AStarArena& arena, cudaStream_t stream) {
  k_arena_init<<<148 * 8, 256, 0, stream>>>(arena);
}

template <uint32_t QPW, uint32_t HS, uint32_t BPS, uint32_t KS, int LAYOUT, bool SIMPLE>
static void launch_astar_t(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                           AStarCounters* counters, const uint32_t* type_raw, cudaStream_t stream) {
  if constexpr ((!SIMPLE || KS != 2) && BPS == 5) {
    // with cost tables / overrides in play (or without the quad specialisation) the kernel does not fit 96
    // registers without spilling (and spill reloads miss the shared-memory-sized L1): four blocks per SM at
    // 113 registers instead
    launch_astar_t<QPW, 48, 4, KS, LAYOUT, SIMPLE>(nav, arena, q, pool, counters, type_raw, stream);
    return;
  } else {
  static_assert(HS >= AT_HEAP_SM_MIN, "a slot's open-list spill is sized for HS >= AT_HEAP_SM_MIN (arena_layout)");
  constexpr uint32_t QPB = QPW * (AT_THREADS / 32);
  const size_t smem = (size_t)HS * QPB * sizeof(uint4);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_astar<QPW, HS, BPS, KS, LAYOUT, SIMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    cudaFuncSetAttribute(k_astar<QPW, HS, BPS, KS, LAYOUT, SIMPLE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  // never more blocks than are resident at once: a block that starts late starts its (long) queries late
  uint32_t slots = arena.n_slots < q.n ? arena.n_slots : q.n;
  const uint32_t resident = (uint32_t)sm_count * BPS * QPB;
  if (slots > resident) slots = resident;
  uint32_t blocks = (slots + QPB - 1) / QPB;
  AStarArena a = arena;
  a.n_slots = slots;
  k_astar<QPW, HS, BPS, KS, LAYOUT, SIMPLE><<<blocks, AT_THREADS, smem, stream>>>(nav, a, q, pool, counters, type_raw);
  }
}

// Launch shapes.  Only half the lanes of a warp own a query (16 per warp): the number of queries
// in flight is bounded by HBM (the per-query arena slot) and by shared memory (the open lists),
// and the search by the latency of one lock-step warp iteration, not by warp or issue slots — a
// warp that waits for the slowest of 16 instead of 32 pops measured 5-14 % faster.
//  * small open lists (mazes: a few dozen open branches): 320 queries per SM = 5 blocks x 4 warps
//    x 16, 40 open-list entries per query in shared memory.  __launch_bounds__(128, 5) holds the
//    kernel to 96 registers without a spill (113 at 4 blocks); the fifth block is worth 7 % on the
//    bench (822 -> 765 ms).  A sixth (32 entries, 80 registers, 68 bytes spilled) is slower than
//    four;
//  * large open lists (open fields: a wavefront of thousands): 128 queries per SM with 96 entries in
//    shared memory.  Every spilled heap level is a dependent HBM round trip and these searches are
//    bound by random-sector HBM traffic, so one more level on chip beats twice the queries in
//    flight (open 256x256 field, 50k queries: 1.47 s -> 1.06 s), and the same 128 queries as 16
//    warps of 8 instead of 8 warps of 16 hide more of those round trips (-> 0.92 s; 160 queries
//    with 80 entries: 0.97 s, 80 with 160: 1.01 s, 64 with 192: 1.15 s).
//  * small batches run at the latency of their longest search, so they get the widest shapes the
//    batch size allows: up to 32 queries per SM -> 4 queries per warp with 384 entries each, up to
//    8 per SM -> one query per warp with 1536 entries (a lone lane is not held back by
//    lock-step neighbours, and an open-field open list fits on chip); up to 128 per SM -> 8 queries
//    per warp with 96 entries.
// (Measured and rejected in round 2, forest layout: evaluating cost / estimate per pushed successor inside the
// push loop instead of for all four edges up front — 5 % less latency for a lone query, but in a loaded warp the
// loop body then runs with ~6 of 16 lanes active and the bench step went 338 -> 349 ms.)
// (Measured and rejected: 32 queries per warp; 384 queries per SM with 28 shared-memory entries;
// 64 / 192 / 384 entries at 192 / 64 / 32 queries per SM; a 4-lanes-per-query variant — lower
// latency per query but issue-bound, no faster when loaded.)
uint32_t astar_resident_queries_per_sm(bool big_heap, uint32_t shape) {
  if (!big_heap && shape == 1) return 6 * 16 * (AT_THREADS / 32);
  if (!big_heap && shape == 2) return 3 * 32 * (AT_THREADS / 32);
  if (!big_heap && shape == 3) return 4 * 32 * (AT_THREADS / 32);
  if (!big_heap && shape == 4) return 7 * 8 * (AT_THREADS / 32);
  return (big_heap ? 2 : 5) * 16 * (AT_THREADS / 32);
}

uint32_t launch_astar_warp(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                           AStarCounters* counters, const uint32_t* type_raw, bool simple, bool pair, uint32_t sm_count,
                           cudaStream_t stream);  // k_astar_warp.cu

uint32_t launch_astar(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                      AStarCounters* counters, const uint32_t* type_raw, bool big_heap, uint32_t flags,
                      uint32_t shape, cudaStream_t stream) {
  if (q.n == 0) return 0;
  const bool k4 = arena.kshift == 2;
  const bool simple = nav.n_types == 1 && !q.override_begin;  // the kernel's SIMPLE flag
#define LMB_ASTAR_LAUNCH_S(QPW, HS, BPS, K, L)                                                             \
  do {                                                                                                    \
    if (simple) launch_astar_t<QPW, HS, BPS, K, L, true>(nav, arena, q, pool, counters, type_raw, stream);  \
    else launch_astar_t<QPW, HS, BPS, K, L, false>(nav, arena, q, pool, counters, type_raw, stream);        \
  } while (0)
#define LMB_ASTAR_LAUNCH_L(QPW, HS, BPS, L)           \
  do {                                               \
    if (k4) LMB_ASTAR_LAUNCH_S(QPW, HS, BPS, 2, L);   \
    else LMB_ASTAR_LAUNCH_S(QPW, HS, BPS, 0, L);      \
  } while (0)
#define LMB_ASTAR_LAUNCH(QPW, HS, BPS)                                  \
  do {                                                                  \
    if (arena.layout == 1) LMB_ASTAR_LAUNCH_L(QPW, HS, BPS, 1);         \
    else if (arena.layout == 2) LMB_ASTAR_LAUNCH_L(QPW, HS, BPS, 2);    \
    else if (arena.layout == 3) LMB_ASTAR_LAUNCH_L(QPW, HS, BPS, 3);    \
    else LMB_ASTAR_LAUNCH_L(QPW, HS, BPS, 0);                           \
  } while (0)
  int dev = 0, sm_count = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  uint32_t kind = 1;
  // LMB_CONFIG_ASTAR_LARGE_SHAPES: tests run the large-batch shapes (the benchmarked ones) at sizes the oracle can check
  const bool force_large = flags & LMB_CONFIG_ASTAR_LARGE_SHAPES;
  // Small batches finish with their longest search: up to 16 queries per SM go to the warp kernels
  // (k_astar_warp.cu: a search warp + a heap warp per query), whose single-search latency is several times lower.
  // (measured, scripts/astar_latency.py, us per explored node of the longest search at 1 / 4 / 8 / 16 queries per
  // SM: maze 0.79 / 0.87 / 0.91 / 1.12-1.16 against 1.05 / 1.08 / 1.12 / 1.35 thread-per-query; open field
  // 1.11 / 1.30 / 1.51 / 2.03 against 1.83 / 1.93 / 1.96 / 3.56.  Without the heap warp — LMB_CONFIG_ASTAR_NO_PAIR
  // — a query's open list has to fit its share of shared memory for the warp kernel to win: up to 4 per SM off a
  // forest.)
  const bool pair = !(flags & LMB_CONFIG_ASTAR_NO_PAIR);
  const uint32_t warp_max = (arena.layout == 3 || pair ? 16u : 4u) * (uint32_t)sm_count;
  const bool warp = (flags & LMB_CONFIG_ASTAR_WARP_ONLY) ||
                    (!force_large && !(flags & LMB_CONFIG_ASTAR_THREAD_ONLY) && q.n <= warp_max);
  if (warp) {
    kind = launch_astar_warp(nav, arena, q, pool, counters, type_raw, simple, pair, (uint32_t)sm_count, stream) == 3u ? 4u : 2u;
  } else if (!force_large && q.n <= 8u * (uint32_t)sm_count)
    LMB_ASTAR_LAUNCH(1, 1536, 2);  // few queries: one per warp, the whole open list on chip
  else if (!force_large && q.n <= 32u * (uint32_t)sm_count)
    LMB_ASTAR_LAUNCH(4, 384, 2);
  else if (!force_large && q.n <= 128u * (uint32_t)sm_count)
    LMB_ASTAR_LAUNCH(8, 96, 4);
  else if (big_heap)
    LMB_ASTAR_LAUNCH(8, 96, 4);
  else if (arena.layout == 3 && k4 && shape == 1)
    LMB_ASTAR_LAUNCH_S(16, 36, 6, 2, 3);  // tuning hooks (lmb_config.astar_shape), forest quad meshes only
  else if (arena.layout == 3 && k4 && shape == 2)
    LMB_ASTAR_LAUNCH_S(32, 32, 3, 2, 3);
  else if (arena.layout == 3 && k4 && shape == 3)
    LMB_ASTAR_LAUNCH_S(32, 26, 4, 2, 3);
  else if (arena.layout == 3 && k4 && shape == 4)
    LMB_ASTAR_LAUNCH_S(8, 56, 7, 2, 3);
  else
    LMB_ASTAR_LAUNCH(16, 40, 5);
#undef LMB_ASTAR_LAUNCH
#undef LMB_ASTAR_LAUNCH_L
#undef LMB_ASTAR_LAUNCH_S
  AStarArena a = arena;
  // corridors of the queries that finished in this launch: raw states -> (node, portal)
  uint32_t fix_blocks = (q.n + 7) / 8;
  if (fix_blocks > 148 * 16) fix_blocks = 148 * 16;
  k_path_fixup<<<fix_blocks, 256, 0, stream>>>(nav, a, q, pool);
  return kind;
}

}  // namespace lmb
