// Kernel 2 — batched A*, one query per warp.
//
// Replaces `pathfinding::find_path` + `ArchipelagoPathProblem` + `astar::find_path`
// (crates/landmass/src/pathfinding.rs:19-382, astar.rs:126-206).
//
// Parity contract: corridors, portal edge indices and `explored_nodes` are bit-exact.
// The open list is therefore NOT a bucketed queue: it is an exact emulation of Rust's
// std `BinaryHeap<Reverse<NodeRef>>` (push = sift_up, pop = sift_down_to_bottom +
// sift_up) with the reference's comparator, including the cost-vs-estimate quirk at
// astar.rs:71.  What the warp parallelises is everything around
// it: the successors of an expansion are evaluated one per lane (one 32-byte EdgeRec
// sector per lane: midpoint, twin state, neighbour type), their `best_estimates`
// probes are issued together, and accepted successors are pushed in ascending edge
// order via a ballot.
//
// State space (pathfinding.rs:55-69): NodeEdge{node, start_edge} = global edge id,
// OffMeshLink(l) = n_edges + l, Start = n_edges + n_links, End = Start + 1.
// `best_estimates` (a HashMap in the reference) is a dense f32 array per warp slot,
// kept at +inf between queries and restored by walking the query's node list.
#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {

namespace {

constexpr float F_INF = __builtin_huge_valf();

struct HeapEntry {
  float cost, estimate;
  uint32_t index;
  uint32_t state;
};

// NodeRef::partial_cmp (astar.rs:67-77) as `a <= b`
__device__ __forceinline__ bool noderef_le(const HeapEntry& a, const HeapEntry& b) {
  if (a.estimate < b.estimate) return true;
  if (a.estimate > b.estimate) return false;
  if (a.estimate == b.estimate) return b.estimate <= a.cost;
  return false;
}
// Reverse<T>::le(x, y) = y.0 <= x.0
__device__ __forceinline__ bool rev_le(const HeapEntry& x, const HeapEntry& y) { return noderef_le(y, x); }

// Open list.  Entries [0, SMEM_HEAP) live in shared memory (one private region per warp),
// the rest spills to the slot's global arena.  Entry = {cost, estimate, node index, state}.
constexpr uint32_t SMEM_HEAP = 384;

template <bool SPILL>
struct Heap {
  uint4* sm;    // [SMEM_HEAP]
  uint4* gm;    // global spill (SPILL only); entry i (>= SMEM_HEAP) at gm[i + 1]
  uint32_t len;
  __device__ __forceinline__ HeapEntry get(uint32_t i) const {
    uint4 v;
    if (SPILL)
      v = i < SMEM_HEAP ? sm[i] : gm[i + 1];
    else
      v = sm[i];
    return {__uint_as_float(v.x), __uint_as_float(v.y), v.z, v.w};
  }
  __device__ __forceinline__ void set(uint32_t i, const HeapEntry& e) {
    uint4 v = make_uint4(__float_as_uint(e.cost), __float_as_uint(e.estimate), e.index, e.state);
    if (!SPILL || i < SMEM_HEAP)
      sm[i] = v;
    else
      gm[i + 1] = v;
  }
  __device__ __forceinline__ void sift_up(uint32_t start, uint32_t pos, const HeapEntry& hole) {
    while (pos > start) {
      uint32_t parent = (pos - 1) >> 1;
      HeapEntry p = get(parent);
      if (rev_le(hole, p)) break;
      set(pos, p);
      pos = parent;
    }
    set(pos, hole);
  }
  __device__ __forceinline__ void push(const HeapEntry& e) {
    uint32_t old_len = len++;
    sift_up(0, old_len, e);
  }
  __device__ __forceinline__ HeapEntry pop() {
    HeapEntry item = get(--len);
    if (len > 0) {
      HeapEntry top = get(0);
      HeapEntry hole = item;  // swap(item, data[0]); the hole element is the old last
      item = top;
      uint32_t end = len, pos = 0, child = 1;
      while (end >= 2 && child <= end - 2) {
        HeapEntry a = get(child), b = get(child + 1);
        bool right = rev_le(a, b);
        set(pos, right ? b : a);
        pos = child + (right ? 1u : 0u);
        child = 2 * pos + 1;
      }
      if (child == end - 1) {
        set(pos, get(child));
        pos = child;
      }
      sift_up(0, pos, hole);
    }
    return item;
  }
};

struct Query {
  uint32_t start_node, end_node;
  V3 start_point, end_point;
};

}  // namespace

constexpr uint32_t ASTAR_WARPS = 8;  // warps per block; 4 blocks/SM -> 32 concurrent queries per SM

// SPILL = false: the open list lives entirely in shared memory; a query whose open list
// outgrows it is reported as status 4 and re-run by the SPILL = true variant.
template <bool SPILL>
__global__ void __launch_bounds__(ASTAR_WARPS * 32, 4)
k_astar(DevNav nav, AStarArena arena, AStarQueries q, PathPool pool, AStarCounters* counters,
        const uint32_t* __restrict__ type_raw) {
  __shared__ float s_cost[ASTAR_WARPS][LMB_MAX_TYPES];
  extern __shared__ uint4 s_heap[];  // [ASTAR_WARPS][SMEM_HEAP]
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_in_block = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (slot >= arena.n_slots) return;
  float* cost_tab = s_cost[warp_in_block];

  uint8_t* slot_base = arena.base + (size_t)slot * arena.slot_stride;
  float* best = reinterpret_cast<float*>(slot_base);
  uint4* nodes = reinterpret_cast<uint4*>(slot_base + arena.nodes_offset);  // cost, state, prev, depth
  Heap<SPILL> heap;
  heap.sm = s_heap + (size_t)warp_in_block * SMEM_HEAP;
  heap.gm = reinterpret_cast<uint4*>(slot_base + arena.heap_offset);

  const uint32_t heap_cap = SPILL ? arena.heap_cap : SMEM_HEAP;
  const uint32_t ST_LINK0 = nav.n_edges;
  const uint32_t ST_START = nav.n_edges + nav.n_links;
  const uint32_t ST_END = ST_START + 1;

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(&counters->queue_head, 1u);
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi >= q.n) break;
    const uint32_t qid = q.list ? q.list[qi] : qi;
    Query qu;
    qu.start_node = q.start_node[qid];
    qu.end_node = q.end_node[qid];
    qu.start_point = {q.start_point[3 * (size_t)qid], q.start_point[3 * (size_t)qid + 1], q.start_point[3 * (size_t)qid + 2]};
    qu.end_point = {q.end_point[3 * (size_t)qid], q.end_point[3 * (size_t)qid + 1], q.end_point[3 * (size_t)qid + 2]};
    const uint32_t owner = q.owner ? q.owner[qid] : qid;

    // --- per-query type costs: override -> archipelago -> 1.0 (pathfinding.rs:73-77) ---
    for (uint32_t t = lane; t < nav.n_types; t += 32) cost_tab[t] = nav.type_cost[t];
    __syncwarp();
    if (q.override_begin) {
      uint32_t ob = q.override_begin[owner], oe = q.override_begin[owner + 1];
      for (uint32_t k = ob + lane; k < oe; k += 32) {
        uint32_t raw = q.override_type[k];
        for (uint32_t t = 0; t < nav.n_types; t++)
          if (type_raw[t] == raw) cost_tab[t] = q.override_cost[k];
      }
      __syncwarp();
    }
    // cheapest_type_index_cost (pathfinding.rs:300-314)
    float cheapest = 1.0f;
    for (uint32_t t = lane; t < nav.n_types; t += 32) {
      float c = cost_tab[t];
      if (nav.type_registered[t] && is_finite(c) && c < cheapest) cheapest = c;
    }
    for (int o = 16; o > 0; o >>= 1) cheapest = fminf(cheapest, __shfl_xor_sync(0xffffffffu, cheapest, o));

    // PermittedAnimationLinks::is_permitted (agent.rs:178-186)
    auto is_kind_permitted = [&](uint32_t kind) {
      if (!q.owner_flags || (q.owner_flags[owner] & LMB_AGENT_PERMIT_ALL_LINKS)) return true;
      if (q.permitted_begin)
        for (uint32_t m = q.permitted_begin[owner]; m < q.permitted_begin[owner + 1]; m++)
          if (q.permitted_kind[m] == kind) return true;
      return false;
    };

    // --- are_nodes_connected (nav_data.rs:1022-1087) ---
    bool connected;
    {
      uint32_t i1 = nav.poly_island[qu.start_node], i2 = nav.poly_island[qu.end_node];
      if (i1 == i2 && nav.poly_region[qu.start_node] == nav.poly_region[qu.end_node]) {
        connected = true;
      } else {
        uint32_t r1 = nav.node_region_number[qu.start_node], r2 = nav.node_region_number[qu.end_node];
        connected = r1 != LMB_NONE && r2 != LMB_NONE && nav.region_root[r1] == nav.region_root[r2];
        if (!connected && r1 != LMB_NONE && r2 != LMB_NONE && nav.n_possible_links) {
          // BFS over region numbers through the permitted PossibleRegionLinks
          // (nav_data.rs:1056-1086).  seen[] / queue[] borrow the (still empty) node arena.
          uint32_t* seen = reinterpret_cast<uint32_t*>(nodes);
          uint32_t* queue = seen + nav.n_region_numbers;
          for (uint32_t r = lane; r < nav.n_region_numbers; r += 32) seen[r] = 0u;
          __syncwarp();
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
          __syncwarp();
        }
      }
    }

    uint32_t explored = 0, success = 0, status = 1;
    uint32_t path_off = 0, path_len = 0;
    uint32_t n_nodes = 0;
    heap.len = 0;

    if (connected) {
      // try_add_node(Start)
      {
        float est = fadd(0.0f, fmul(distance(qu.start_point, qu.end_point), cheapest));
        if (!(F_INF <= est)) {
          best[ST_START] = est;
          nodes[0] = make_uint4(__float_as_uint(0.0f), ST_START, LMB_NONE, 0u);
          n_nodes = 1;
          heap.push({0.0f, est, 0u, ST_START});
        }
      }
      uint32_t goal = LMB_NONE;
      while (heap.len > 0) {
        HeapEntry cur = heap.pop();
        const uint32_t s = cur.state;
        // Issue every load the expansion needs before the first use: best[s] (stale test) and
        // the node record (depth) usually come from HBM, the edge records from L2.
        const float best_s = best[s];
        const uint32_t cur_depth = nodes[cur.index].w;
        EdgeRec self_rec;
        if (s < ST_LINK0) self_rec = nav.edges[s];

        // resolve (node, point, ignore step) — pathfinding.rs:93-143
        uint32_t poly, first, ne, cur_type, has_links;
        uint32_t ignore_edge = LMB_NONE, ignore_link = LMB_NONE;
        V3 point;
        if (s < ST_LINK0) {
          poly = self_rec.poly;
          first = self_rec.first_edge;
          ne = self_rec.info & 0xffu;
          cur_type = (self_rec.info >> 8) & 0xffu;
          has_links = (self_rec.info >> 24) & 1u;
          point = {self_rec.mx, self_rec.my, self_rec.mz};
          ignore_edge = s;
        } else if (s == ST_END) {
          poly = first = ne = cur_type = has_links = 0;
          point = {0.0f, 0.0f, 0.0f};
        } else {
          if (s == ST_START) {
            poly = qu.start_node;
            point = qu.start_point;
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
          first = nav.poly_edge_begin[poly];
          ne = nav.poly_edge_begin[poly + 1] - first;
          cur_type = nav.poly_type_dense[poly];
          has_links = nav.node_link_begin[poly + 1] > nav.node_link_begin[poly];
        }
        // first chunk of successor edge records, speculatively (before the stale test resolves)
        EdgeRec succ_rec;
        succ_rec.twin = LMB_NONE;
        if (s != ST_END && poly != qu.end_node && lane < ne) succ_rec = nav.edges[first + lane];

        if (best_s < cur.estimate) continue;  // stale (astar.rs:172-176)
        explored++;
        if (s == ST_END) {
          goal = cur.index;
          break;
        }
        const float cur_cost_so_far = cur.cost;
        const uint32_t depth1 = cur_depth + 1;
        const float cur_type_cost = cost_tab[cur_type];

        bool overflow = false;
        if (poly == qu.end_node) {
          // single successor: GoToEnd (pathfinding.rs:152-155); h(End) = 0
          float c = fmul(distance(point, qu.end_point), cur_type_cost);
          float ncost = fadd(cur_cost_so_far, c);
          float est = fadd(ncost, 0.0f);
          if (!(best[ST_END] <= est)) {
            if (n_nodes >= arena.node_cap || heap.len >= heap_cap) {
              overflow = true;
            } else {
              best[ST_END] = est;
              nodes[n_nodes] = make_uint4(__float_as_uint(ncost), ST_END, cur.index, depth1);
              heap.push({ncost, est, n_nodes, ST_END});
              n_nodes++;
            }
          }
        } else {
          // node connections, ascending edge index, one edge per lane
          for (uint32_t e0 = 0; e0 < ne && !overflow; e0 += 32) {
            uint32_t e = e0 + lane;
            bool accept = false;
            float ncost = 0.0f, est = 0.0f;
            uint32_t nstate = LMB_NONE;
            if (e < ne) {
              const EdgeRec r = e0 == 0 ? succ_rec : nav.edges[first + e];
              if (r.twin != LMB_NONE && first + e != ignore_edge &&
                  is_finite(cost_tab[(r.info >> 16) & 0xffu])) {
                V3 mid{r.mx, r.my, r.mz};
                float c = fmul(distance(point, mid), cur_type_cost);
                ncost = fadd(cur_cost_so_far, c);
                est = fadd(ncost, fmul(distance(mid, qu.end_point), cheapest));
                nstate = r.twin;
                accept = !(best[nstate] <= est);
              }
            }
            uint32_t mask = __ballot_sync(0xffffffffu, accept);
            while (mask) {
              int src = __ffs(mask) - 1;
              mask &= mask - 1;
              float c2 = __shfl_sync(0xffffffffu, ncost, src);
              float e2 = __shfl_sync(0xffffffffu, est, src);
              uint32_t s2 = __shfl_sync(0xffffffffu, nstate, src);
              if (n_nodes >= arena.node_cap || heap.len >= heap_cap) {
                overflow = true;
                break;
              }
              best[s2] = e2;
              nodes[n_nodes] = make_uint4(__float_as_uint(c2), s2, cur.index, depth1);
              heap.push({c2, e2, n_nodes, s2});
              n_nodes++;
            }
          }
          // off-mesh links in ascending link index (canonical; pathfinding.rs:196-229)
          if (has_links && !overflow) {
            for (uint32_t k = nav.node_link_begin[poly]; k < nav.node_link_begin[poly + 1]; k++) {
              uint32_t l = nav.node_links[k];
              if (l == ignore_link) continue;
              if (!is_finite(cost_tab[nav.link_dst_type_dense[l]])) continue;
              float link_cost = 0.0f;
              if (nav.link_kind[l] == LMB_LINK_ANIMATION) {
                if (!is_kind_permitted(nav.link_anim_kind[l])) continue;
                link_cost = nav.link_cost[l];
              }
              const float* pp = nav.link_portal + 6 * (size_t)l;
              V3 pm = midpoint(V3{pp[0], pp[1], pp[2]}, V3{pp[3], pp[4], pp[5]});
              float c = fadd(fmul(distance(point, pm), cur_type_cost), link_cost);
              float ncost = fadd(cur_cost_so_far, c);
              // heuristic of OffMeshLink state: portal (boundary) / destination portal (animation)
              V3 hp = pm;
              if (nav.link_kind[l] != LMB_LINK_BOUNDARY) {
                const float* dp = nav.link_dst_portal + 6 * (size_t)l;
                hp = midpoint(V3{dp[0], dp[1], dp[2]}, V3{dp[3], dp[4], dp[5]});
              }
              float est = fadd(ncost, fmul(distance(hp, qu.end_point), cheapest));
              uint32_t s2 = ST_LINK0 + l;
              if (best[s2] <= est) continue;
              if (n_nodes >= arena.node_cap || heap.len >= heap_cap) {
                overflow = true;
                break;
              }
              best[s2] = est;
              nodes[n_nodes] = make_uint4(__float_as_uint(ncost), s2, cur.index, depth1);
              heap.push({ncost, est, n_nodes, s2});
              n_nodes++;
            }
          }
        }
        if (overflow) {
          status = (!SPILL && n_nodes < arena.node_cap) ? 4 : 2;
          break;
        }
      }

      // --- recover the path (astar.rs:86-105, pathfinding.rs:324-381) into the pool ---
      if (status == 1 && goal != LMB_NONE) {
        const uint32_t L = nodes[goal].w;  // number of corridor nodes
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(pool.cursor, (unsigned long long)L);
        off = __shfl_sync(0xffffffffu, off, 0);
        if (off + L > pool.capacity) {
          status = 3;
        } else {
          uint32_t i = goal;
          uint32_t next_state = ST_END;
          for (uint32_t pos = L; pos-- > 0;) {
            i = nodes[i].z;
            uint32_t st = nodes[i].y;
            uint32_t node_id;
            if (st < ST_LINK0)
              node_id = nav.edges[st].poly;
            else if (st == ST_START)
              node_id = qu.start_node;
            else
              node_id = nav.link_dst_node[st - ST_LINK0];
            uint32_t portal;
            if (next_state == ST_END) {
              portal = LMB_NONE;
            } else if (next_state < ST_LINK0) {
              uint32_t tw = nav.edges[next_state].twin;  // the edge inside `node_id`
              portal = tw - nav.edges[tw].first_edge;
            } else {
              portal = LMB_PORTAL_LINK | (next_state - ST_LINK0);
            }
            if (lane == 0) pool.entries[off + pos] = make_uint2(node_id, portal);
            next_state = st;
          }
          success = 1;
          path_off = (uint32_t)off;
          path_len = L;
        }
      }

      // --- restore best_estimates to +inf for the next query of this slot ---
      __syncwarp();
      for (uint32_t k = lane; k < n_nodes; k += 32) best[nodes[k].y] = F_INF;
      __syncwarp();
    }

    if (lane == 0) {
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
      } else if (status == 4) {
        atomicAdd(&counters->n_retry_spill, 1u);
      } else {
        atomicAdd(&counters->n_retry_pool, 1u);
      }
    }
  }
}

__global__ void k_arena_init(AStarArena arena) {
  size_t total = (size_t)arena.n_slots * arena.n_states;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t slot = i / arena.n_states, k = i % arena.n_states;
    reinterpret_cast<float*>(arena.base + slot * arena.slot_stride)[k] = F_INF;
  }
}

void launch_arena_init(const AStarArena& arena, cudaStream_t stream) {
  k_arena_init<<<148 * 8, 256, 0, stream>>>(arena);
}

void launch_astar(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                  AStarCounters* counters, const uint32_t* type_raw, bool spill, cudaStream_t stream) {
  if (q.n == 0) return;
  const uint32_t warps_per_block = ASTAR_WARPS;
  const size_t smem = (size_t)ASTAR_WARPS * SMEM_HEAP * sizeof(uint4);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_astar<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_astar<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_astar<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_astar<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  uint32_t slots = arena.n_slots < q.n ? arena.n_slots : q.n;
  uint32_t blocks = (slots + warps_per_block - 1) / warps_per_block;
  AStarArena a = arena;
  a.n_slots = blocks * warps_per_block <= arena.n_slots ? blocks * warps_per_block : arena.n_slots;
  if (spill)
    k_astar<true><<<blocks, warps_per_block * 32, smem, stream>>>(nav, a, q, pool, counters, type_raw);
  else
    k_astar<false><<<blocks, warps_per_block * 32, smem, stream>>>(nav, a, q, pool, counters, type_raw);
}

}  // namespace lmb
