// Kernel 2 (quad variant) — batched A*, FOUR lanes per query, eight queries per warp.
//
// Same algorithm, state ids, arena layout and outputs as k_astar.cu (see there for the parity
// contract and the memory layout).  What changes is the mapping: the number of concurrent
// queries is bounded by HBM (one dense `best_estimates` array per query), which leaves most of
// the SM's issue slots idle with one lane per query, and the time of a batch is bounded below
// by its longest query = (pops + corridor length) x the latency of one iteration.  A quad
// shortens that iteration:
//   * lane j of the quad loads, probes and evaluates edge j of the expanded polygon (the four
//     records of a padded polygon are one 128-byte line), so an expansion costs one record
//     load, one `best` probe and two square roots per lane instead of four of each in series;
//   * the open list (exact BinaryHeap emulation) is touched by the quad's leader lane only; the
//     popped entry and the accepted successors are exchanged with warp shuffles;
//   * eight instead of thirty-two searches run in lock-step, so a warp iteration waits for the
//     slowest of 8 pops / memory round trips, not of 32.
// Every decision (stale test, accept test, push order = ascending edge index) is the one the
// thread-per-query kernel makes, so results are identical by construction.
#include "k_astar_common.cuh"

namespace lmb {

using namespace astar_detail;

constexpr uint32_t AQ_QPW = 8;                       // queries per warp
constexpr uint32_t AQ_QPB = AQ_QPW * (AT_THREADS / 32);  // queries per block

__global__ void __launch_bounds__(AT_THREADS, AQ_BLOCKS_PER_SM)
k_astar_quad(DevNav nav, AStarArena arena, AStarQueries q, PathPool pool, AStarCounters* counters,
             const uint32_t* __restrict__ type_raw) {
  extern __shared__ uint4 s_heap[];  // [AT_HEAP_SM][AQ_QPB] uint4, then the same shape of u32 depths
  __shared__ float s_cost[LMB_MAX_TYPES];
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t ql = lane & 3;           // lane within the quad
  const uint32_t qbase = lane & ~3u;      // first lane of the quad
  const bool leader = ql == 0;
  const uint32_t qib = (tid >> 5) * AQ_QPW + (lane >> 2);  // query index in block
  const uint32_t slot = blockIdx.x * AQ_QPB + qib;
  for (uint32_t t = tid; t < nav.n_types; t += AT_THREADS) s_cost[t] = nav.type_cost[t];
  __syncthreads();

  const uint32_t slot_c = slot < arena.n_slots ? slot : 0;  // out-of-range quads never touch memory
  uint8_t* const slot_base = arena.base + (size_t)slot_c * arena.slot_stride;
  float* const best = reinterpret_cast<float*>(slot_base);
  uint2* const nodes = reinterpret_cast<uint2*>(slot_base + arena.nodes_offset);  // {state, prev}
  Heap<AQ_QPB> heap;  // used by the leader; `len` is mirrored in every lane of the quad
  heap.sm = s_heap + qib;
  heap.smd = reinterpret_cast<uint32_t*>(s_heap + (size_t)AT_HEAP_SM * AQ_QPB) + qib;
  heap.gm = reinterpret_cast<uint4*>(slot_base + arena.heap_offset);
  heap.gmd = reinterpret_cast<uint32_t*>(slot_base + arena.heapd_offset);
  heap.len = 0;

  const EdgeRec* __restrict__ recs = nav.edges;
  const uint32_t heap_cap = arena.heap_cap;
  const uint32_t node_cap = arena.node_cap;
  const uint32_t kshift = arena.kshift;
  const uint32_t kmask = (1u << kshift) - 1u;
  const uint32_t ST_LINK0 = arena.n_ids;
  const uint32_t ST_START = arena.n_ids + nav.n_links;
  const uint32_t ST_END = ST_START + 1;

  // ---- quad state (identical in the four lanes unless noted) ----
  uint32_t phase = slot < arena.n_slots ? PH_IDLE : PH_EXIT;
  uint32_t qid = 0, owner = 0, start_node = 0, end_node = 0;
  V3 end_point{0.0f, 0.0f, 0.0f};
  float cheapest = 1.0f;
  uint32_t ov_begin = 0, ov_end = 0;  // this query's cost overrides
  bool permit_all = true;
  uint32_t n_nodes = 0, explored = 0, status = 1, success = 0;
  uint32_t path_off = 0, path_len = 0;
  uint32_t w_node = 0, w_next = 0, w_pos = 0;  // walk

  // type cost: override -> archipelago -> 1.0 (pathfinding.rs:73-77)
  auto cost_of = [&](uint32_t t) -> float {
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
    if (permit_all) return true;
    if (q.permitted_begin)
      for (uint32_t m = q.permitted_begin[owner]; m < q.permitted_begin[owner + 1]; m++)
        if (q.permitted_kind[m] == kind) return true;
    return false;
  };
  auto poly_span = [&](uint32_t poly, uint32_t& first, uint32_t& cnt) {
    if (kshift) {
      first = poly << kshift;
      cnt = kmask + 1u;
    } else {
      first = nav.poly_edge_begin[poly];
      cnt = nav.poly_edge_begin[poly + 1] - first;
    }
  };
  // try_add_node's tail: record the node and push it (stores and heap work by the leader only)
  auto add_node = [&](uint32_t s2, float ncost, float est, uint32_t prev, uint32_t depth) {
    if (leader) {
      best[s2] = est;
      nodes[n_nodes] = make_uint2(s2, prev);
      heap.push({ncost, est, n_nodes, s2, depth});
    } else {
      heap.len++;
    }
    n_nodes++;
  };

  for (;;) {
    // =========================== IDLE: fetch and set up a query ===========================
    if (phase == PH_IDLE) {
      uint32_t qi = 0;
      if (leader) qi = atomicAdd(&counters->queue_head, 1u);
      qi = __shfl_sync(0xfu << qbase, qi, qbase);
      if (qi >= q.n) {
        phase = PH_EXIT;
      } else {
        qid = q.list ? q.list[qi] : qi;
        start_node = q.start_node[qid];
        end_node = q.end_node[qid];
        const V3 start_point{q.start_point[3 * (size_t)qid], q.start_point[3 * (size_t)qid + 1],
                             q.start_point[3 * (size_t)qid + 2]};
        end_point = {q.end_point[3 * (size_t)qid], q.end_point[3 * (size_t)qid + 1],
                     q.end_point[3 * (size_t)qid + 2]};
        owner = q.owner ? q.owner[qid] : qid;
        ov_begin = ov_end = 0;
        if (q.override_begin) {
          ov_begin = q.override_begin[owner];
          ov_end = q.override_begin[owner + 1];
        }
        permit_all = !q.owner_flags || (q.owner_flags[owner] & LMB_AGENT_PERMIT_ALL_LINKS);
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
              // (nav_data.rs:1056-1086), by the leader; seen[] / queue[] borrow the (still
              // empty) node list.
              uint32_t found = 0;
              if (leader) {
                uint32_t* seen = reinterpret_cast<uint32_t*>(nodes);
                uint32_t* queue = seen + nav.n_region_numbers;
                for (uint32_t r = 0; r < nav.n_region_numbers; r++) seen[r] = 0u;
                uint32_t head = 0, tail = 0;
                seen[r1] = 1u;
                queue[tail++] = r1;
                while (head < tail) {
                  uint32_t r = queue[head++];
                  if (r == r2) {
                    found = 1;
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
              connected = __shfl_sync(0xfu << qbase, found, qbase) != 0;
            }
          }
        }
        explored = 0;
        success = 0;
        status = 1;
        path_off = path_len = 0;
        n_nodes = 0;
        heap.len = 0;
        phase = PH_FINISH;
        if (connected) {
          // try_add_node(Start)
          float est = fadd(0.0f, fmul(distance(start_point, end_point), cheapest));
          if (!(F_INF <= est)) {
            add_node(ST_START, 0.0f, est, LMB_NONE, 0u);
            phase = PH_SEARCH;
          }
        }
      }
    }
    // stores of this iteration's set-up / previous iteration's pushes (leader) are ordered
    // before the probes the other lanes issue below
    __syncwarp();
    if (__all_sync(FULL, phase == PH_EXIT)) break;

    // ============ WALK (part 1): issue this iteration's back-pointer load early ============
    uint2 w_rec = make_uint2(0u, 0u);
    if (phase == PH_WALK) w_rec = nodes[w_node];

    // =============================== SEARCH: one pop ===============================
    if (phase == PH_SEARCH) {
      if (heap.len == 0) {
        phase = PH_FINISH;  // no path
      } else {
        const uint32_t qmask = 0xfu << qbase;
        // ---- the leader peeks the root state; every lane issues its share of the expansion's
        //      loads (best[s] for the stale test, one edge record each) before the pop ----
        uint32_t s = 0;
        if (leader) s = heap.sm[0].w;
        s = __shfl_sync(qmask, s, qbase);
        const float best_s = __ldcg(best + s);
        const bool is_edge = s < ST_LINK0;
        EdgeRec r;
        r.twin = LMB_NONE;
        r.info = 0;
        r.mx = r.my = r.mz = 0.0f;
        r.poly = r.first_edge = 0;
        uint32_t first = 0, cnt = 0;
        const bool preloaded = is_edge && kshift;
        if (preloaded) {
          first = s & ~kmask;
          cnt = kmask + 1u;
          r = ld_rec(recs + first + ql);
        }
        EdgeRec self;
        self.poly = self.first_edge = self.info = 0;
        self.mx = self.my = self.mz = 0.0f;
        if (is_edge && (!kshift || (s & kmask) >= 4)) self = ld_rec(recs + s);

        // ---- pop (leader), broadcast the entry ----
        HeapEntry cur{0.0f, 0.0f, 0u, 0u, 0u};
        if (leader)
          cur = heap.pop();
        else
          heap.len--;
        cur.cost = __shfl_sync(qmask, cur.cost, qbase);
        cur.estimate = __shfl_sync(qmask, cur.estimate, qbase);
        cur.index = __shfl_sync(qmask, cur.index, qbase);
        cur.depth = __shfl_sync(qmask, cur.depth, qbase);

        // ---- resolve (node, point, ignore step) — pathfinding.rs:93-143 ----
        uint32_t poly = 0, cur_type = 0, has_links = 0;
        uint32_t ignore_edge = LMB_NONE, ignore_link = LMB_NONE;
        V3 point{0.0f, 0.0f, 0.0f};
        if (is_edge) {
          if (kshift && (s & kmask) < 4) {
            // the state's own record is one of the four just loaded
            const uint32_t src = qbase + (s & 3u);
            self.mx = __shfl_sync(qmask, r.mx, src);
            self.my = __shfl_sync(qmask, r.my, src);
            self.mz = __shfl_sync(qmask, r.mz, src);
            self.poly = __shfl_sync(qmask, r.poly, src);
            self.info = __shfl_sync(qmask, r.info, src);
          }
          if (!kshift) {
            first = self.first_edge;
            cnt = self.info & 0xffu;
          }
          poly = self.poly;
          cur_type = (self.info >> 8) & 0xffu;
          has_links = (self.info >> 24) & 1u;
          point = {self.mx, self.my, self.mz};
          ignore_edge = s;
        } else if (s != ST_END) {
          if (s == ST_START) {
            poly = start_node;
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
        const uint32_t depth1 = cur.depth + 1;
        const float cur_type_cost = cost_of(cur_type);
        const bool to_end = poly == end_node;  // single successor GoToEnd (pathfinding.rs:152-155)

        // ---- node connections in ascending edge index, four records (one per lane) at a time:
        //      record (L2) -> `best` probe (HBM, speculative: a load) -> cost/estimate ->
        //      stale test -> accepted successors pushed in lane (= edge) order ----
        bool overflow = false, live = true;
        for (uint32_t e0 = 0; (e0 < cnt && live && !overflow) || e0 == 0; e0 += 4) {
          if (e0 > 0 || !preloaded) {
            if (e0 + ql < cnt && s != ST_END)
              r = ld_rec(recs + first + e0 + ql);
            else
              r.twin = LMB_NONE;
          }
          bool cand = r.twin != LMB_NONE && first + e0 + ql != ignore_edge && !to_end;
          const float probe = cand ? __ldcg(best + r.twin) : 0.0f;
          if (!is_finite(cost_of((r.info >> 16) & 0xffu))) cand = false;
          const V3 mid{r.mx, r.my, r.mz};
          const float c = fmul(distance(point, mid), cur_type_cost);
          const float nc = fadd(cur_cost, c);
          const float es = fadd(nc, fmul(distance(mid, end_point), cheapest));
          if (e0 == 0) {
            if (best_s < cur.estimate) {  // stale (astar.rs:172-176)
              live = false;
              break;
            }
            explored++;
            if (s == ST_END) {
              // goal: allocate the corridor and start the walk (astar.rs:86-105)
              const uint32_t L = cur.depth;  // number of corridor nodes (>= 1: Start)
              unsigned long long off = 0;
              if (leader) off = atomicAdd(pool.cursor, (unsigned long long)L);
              off = __shfl_sync(qmask, off, qbase);
              if (off + L > pool.capacity) {
                status = 3;
                phase = PH_FINISH;
              } else {
                success = 1;
                path_off = (uint32_t)off;
                path_len = L;
                w_node = cur.index;  // the End node; its record is loaded next iteration
                w_next = ST_END;
                w_pos = L;
                phase = PH_WALK;
              }
              live = false;
              break;
            }
            if (to_end) {
              // h(End) = 0
              float c2 = fmul(distance(point, end_point), cur_type_cost);
              float ncost = fadd(cur_cost, c2);
              float est = fadd(ncost, 0.0f);
              if (!(__ldcg(best + ST_END) <= est)) {
                if (n_nodes >= node_cap || heap.len >= heap_cap)
                  overflow = true;
                else
                  add_node(ST_END, ncost, est, cur.index, depth1);
              }
              live = false;  // no other successor
              break;
            }
          }
          const bool accept = cand && !(probe <= es);
          uint32_t acc = (__ballot_sync(qmask, accept) >> qbase) & 0xfu;
          while (acc) {
            const uint32_t j = __ffs(acc) - 1;
            acc &= acc - 1;
            const uint32_t s2 = __shfl_sync(qmask, r.twin, qbase + j);
            const float est = __shfl_sync(qmask, es, qbase + j);
            const float ncost = __shfl_sync(qmask, nc, qbase + j);
            if (n_nodes >= node_cap || heap.len >= heap_cap) {
              overflow = true;
              break;
            }
            add_node(s2, ncost, est, cur.index, depth1);
          }
        }
        // ---- off-mesh links in ascending link index (canonical; pathfinding.rs:196-229);
        //      evaluated by every lane (same addresses), stored / pushed by the leader ----
        if (live && has_links && !overflow) {
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
            // the probe is read by the leader and broadcast: a link state pushed by an earlier
            // link of this same loop is impossible (distinct links = distinct states), but the
            // value must be the same in the four lanes
            float pb = 0.0f;
            if (leader) pb = __ldcg(best + s2);
            pb = __shfl_sync(qmask, pb, qbase);
            if (pb <= est) continue;
            if (n_nodes >= node_cap || heap.len >= heap_cap) {
              overflow = true;
              break;
            }
            add_node(s2, ncost, est, cur.index, depth1);
          }
        }
        if (overflow) {
          status = 2;
          phase = PH_FINISH;
        }
      }
    } else if (phase == PH_WALK) {
      // ============ WALK (part 2): consume the back-pointer loaded above ============
      // (astar.rs:86-105) one node per iteration, written backwards as (state, next state)
      if (leader && w_rec.x != ST_END)
        pool.entries[(size_t)path_off + w_pos] = make_uint2(w_rec.x, w_next);
      w_next = w_rec.x;
      if (w_rec.y == LMB_NONE) {
        phase = PH_FINISH;  // wrote the Start node (position 0)
      } else {
        w_node = w_rec.y;
        w_pos--;
      }
    }

    // ====== FINISH: restore best_estimates to +inf (warp-cooperative), publish results ======
    uint32_t fin = __ballot_sync(FULL, phase == PH_FINISH && leader);
    while (fin) {
      const int src = __ffs(fin) - 1;
      fin &= fin - 1;
      const uint32_t nn = __shfl_sync(FULL, n_nodes, src);
      if (nn == 0) continue;
      const uint32_t src_slot = __shfl_sync(FULL, slot, src);
      uint8_t* base = arena.base + (size_t)src_slot * arena.slot_stride;
      float* b = reinterpret_cast<float*>(base);
      if ((size_t)nn * 16 >= arena.n_states) {
        // dense restore: coalesced 16-byte stores over the whole array (padded to 16 B)
        float4* b4 = reinterpret_cast<float4*>(b);
        const uint32_t n4 = (arena.n_states + 3) >> 2;
        const float4 inf4 = make_float4(F_INF, F_INF, F_INF, F_INF);
        for (uint32_t k = lane; k < n4; k += 32) b4[k] = inf4;
      } else {
        const uint2* nd = reinterpret_cast<const uint2*>(base + arena.nodes_offset);
        for (uint32_t k = lane; k < nn; k += 32) b[nd[k].x] = F_INF;
      }
    }
    __syncwarp();
    if (phase == PH_FINISH) {
      if (leader) {
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
          }
          atomicAdd(&counters->n_done, 1u);
          atomicAdd(&counters->explored_total, (unsigned long long)explored);
          atomicAdd(&counters->path_nodes_total, (unsigned long long)path_len);
        } else if (status == 2) {
          atomicAdd(&counters->n_retry_arena, 1u);
        } else {
          atomicAdd(&counters->n_retry_pool, 1u);
        }
      }
      n_nodes = 0;
      phase = PH_IDLE;
    }
  }
}

void launch_astar_quad(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                       AStarCounters* counters, const uint32_t* type_raw, cudaStream_t stream) {
  const size_t smem = (size_t)AT_HEAP_SM * AQ_QPB * (sizeof(uint4) + sizeof(uint32_t));
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_astar_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_astar_quad, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  uint32_t slots = arena.n_slots < q.n ? arena.n_slots : q.n;
  uint32_t blocks = (slots + AQ_QPB - 1) / AQ_QPB;
  AStarArena a = arena;
  a.n_slots = slots;
  k_astar_quad<<<blocks, AT_THREADS, smem, stream>>>(nav, a, q, pool, counters, type_raw);
}

}  // namespace lmb
