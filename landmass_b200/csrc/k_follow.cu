// Kernel 3 — repath decision, funnel string-pulling, reached-target test, desired move.
//
// Replaces phases B (decision part) and C of `Archipelago::update`
// (crates/landmass/src/lib.rs:343-523): `does_agent_need_repath` (agent.rs:425-475),
// `Path::{find_index_of_node, find_index_of_node_rev}` (path.rs:403-446),
// `Path::find_next_point_in_straight_path` (path.rs:227-368),
// `Agent::has_reached_target` (agent.rs:330-404) and the desired-move computation
// (lib.rs:470-521).
//
// Device path representation: one flat list of (node id, portal code) per agent slot in
// a bump-allocated pool.  The flat index is order-isomorphic to the reference's
// `PathIndex {segment_index, portal_index}` and `PathIndex::next` is `+1`
// (path.rs:153-176), see DESIGN.md.
#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {

namespace {

struct PathView {
  const uint2* e;
  uint32_t len;
};

struct Portal {
  bool walkable;
  V3 a, b;          // Walkable(left, right) or the animation link's start portal
  V3 end_a, end_b;  // animation link destination portal
  uint32_t link_id, start_node, end_node;
};

// Path::get_portal_endpoints (path.rs:81-129, 204-218) at flat index i (< len-1 ... any entry
// with a portal code other than LMB_NONE).
__device__ Portal get_portal_endpoints(const DevNav& nav, const PathView& path, uint32_t i) {
  Portal p;
  uint2 ent = path.e[i];
  if (ent.y & LMB_PORTAL_LINK) {
    uint32_t l = ent.y & 0x7fffffffu;
    const float* pp = nav.link_portal + 6 * (size_t)l;
    p.a = {pp[0], pp[1], pp[2]};
    p.b = {pp[3], pp[4], pp[5]};
    if (nav.link_kind[l] == LMB_LINK_BOUNDARY) {
      p.walkable = true;
    } else {
      p.walkable = false;
      const float* dp = nav.link_dst_portal + 6 * (size_t)l;
      p.end_a = {dp[0], dp[1], dp[2]};
      p.end_b = {dp[3], dp[4], dp[5]};
      p.link_id = nav.link_anim_id[l];
      p.start_node = ent.x;
      p.end_node = nav.link_dst_node[l];
    }
  } else {
    // get_edge_indices (nav_mesh.rs:945-950): left = vertex[edge+1], right = vertex[edge];
    // world vertices are transform.apply(local) precomputed at upload (same function, same bits).
    // (one 32-byte record of the per-edge portal table: the world vertices themselves, copied at upload)
    const uint32_t rec = nav.rec_kshift ? ((ent.x << nav.rec_kshift) | ent.y) : nav.poly_edge_begin[ent.x] + ent.y;
    const float4 l = __ldg(nav.portal_lr + 2 * (size_t)rec), r = __ldg(nav.portal_lr + 2 * (size_t)rec + 1);
    p.walkable = true;
    p.a = {l.x, l.y, l.z};
    p.b = {r.x, r.y, r.z};
  }
  return p;
}

struct Step {
  bool is_link;
  V3 point, end_point;
  uint32_t link_id, start_node, end_node;
};

__device__ __forceinline__ float triangle_area_2(V3 p0, V3 p1, V3 p2) {
  return perp_dot(xy(p1) - xy(p0), xy(p2) - xy(p0));
}

// geometry.rs:176-189
__device__ void project_point_to_line_segment(V3 point, V3 start, V3 end, V3& out, float& fraction) {
  V3 dir = end - start;
  float l2 = length_squared(dir);
  if (l2 == 0.0f) {
    out = start;
    fraction = 0.0f;
    return;
  }
  V3 rel = point - start;
  fraction = clampf(fdiv(dot(rel, dir), l2), 0.0f, 1.0f);
  out = dir * fraction + start;
}

__device__ Step make_link_step(const Portal& p, V3 apex) {
  Step st;
  float fraction;
  project_point_to_line_segment(apex, p.a, p.b, st.point, fraction);
  st.is_link = true;
  st.end_point = lerp(p.end_a, p.end_b, fraction);
  st.link_id = p.link_id;
  st.start_node = p.start_node;
  st.end_node = p.end_node;
  return st;
}

// Path::find_next_point_in_straight_path (path.rs:227-368); indices are flat.
__device__ uint32_t find_next_point_in_straight_path(const DevNav& nav, const PathView& path,
                                                     uint32_t start_index, V3 start_point,
                                                     uint32_t end_index, V3 end_point, Step& out) {
  const V3 apex = start_point;
  uint32_t left_index = start_index, right_index = start_index;
  bool end_is_link = false;
  Step end_link;
  V3 current_left, current_right;
  if (start_index == end_index) {
    current_left = current_right = end_point;
  } else {
    Portal p = get_portal_endpoints(nav, path, start_index);
    if (p.walkable) {
      current_left = p.a;
      current_right = p.b;
    } else {
      out = make_link_step(p, apex);
      return start_index + 1;
    }
  }
  uint32_t portal_index = start_index + 1;
  while (portal_index <= end_index) {
    V3 portal_left, portal_right;
    if (portal_index == end_index) {
      portal_left = portal_right = end_point;
    } else {
      Portal p = get_portal_endpoints(nav, path, portal_index);
      if (p.walkable) {
        portal_left = p.a;
        portal_right = p.b;
      } else {
        end_link = make_link_step(p, apex);
        end_is_link = true;
        end_index = portal_index;
        portal_left = portal_right = end_link.point;
      }
    }
    if (triangle_area_2(apex, current_right, portal_right) >= 0.0f) {
      if (triangle_area_2(apex, current_left, portal_right) <= 0.0f) {
        right_index = portal_index;
        current_right = portal_right;
      } else {
        out.is_link = false;
        out.point = current_left;
        return left_index;
      }
    }
    if (triangle_area_2(apex, current_left, portal_left) <= 0.0f) {
      if (triangle_area_2(apex, current_right, portal_left) >= 0.0f) {
        left_index = portal_index;
        current_left = portal_left;
      } else {
        out.is_link = false;
        out.point = current_right;
        return right_index;
      }
    }
    portal_index = portal_index + 1;
  }
  if (!end_is_link) {
    out.is_link = false;
    out.point = end_point;
    return end_index;
  }
  out = end_link;
  return portal_index;
}

__device__ __forceinline__ float reach_distance_or_radius(float reach_distance, float radius) {
  return reach_distance >= 0.0f ? reach_distance : radius;
}

// Agent::has_reached_target (agent.rs:330-404)
__device__ bool has_reached_target(const DevNav& nav, const PathView& path, uint32_t condition,
                                   float dist, V3 sampled_point, uint32_t next_index, const Step& next,
                                   uint32_t target_index, V3 target_point) {
  float dd = fmul(dist, dist);
  if (condition == LMB_REACH_DISTANCE) return distance_squared(sampled_point, target_point) < dd;
  if (condition == LMB_REACH_VISIBLE_AT_DISTANCE) {
    if (next.is_link) return false;
    return next_index == target_index && distance_squared(sampled_point, next.point) < dd;
  }
  if (distance_squared(sampled_point, target_point) > dd) return false;
  if (next_index == target_index) return true;
  if (next.is_link) return false;
  float straight = distance(sampled_point, next.point);
  uint32_t cur_index = next_index;
  V3 cur_point = next.point;
  while (cur_index != target_index && straight < dist) {
    Step nw;
    uint32_t ni = find_next_point_in_straight_path(nav, path, cur_index, cur_point, target_index,
                                                   target_point, nw);
    if (nw.is_link) return false;
    straight = fadd(straight, distance(cur_point, nw.point));
    cur_index = ni;
    cur_point = nw.point;
  }
  return straight < dist;
}

// ValidNavigationMesh::sample_point_on_node (nav_mesh.rs:742-814); island-local point
__device__ V3 sample_point_on_node(const DevNav& nav, V3 point, uint32_t node) {
  const DevIsland& is = nav.islands[nav.poly_island[node]];
  const float EPS = -1e-5f;
  uint32_t ntri, eb = 0, bv = 0, bt = 0;
  if (is.has_height) {
    const uint32_t* hp = nav.hpoly + (size_t)node * 4;
    bv = hp[0];
    bt = hp[2];
    ntri = hp[3];
  } else {
    eb = nav.poly_edge_begin[node];
    ntri = nav.poly_edge_begin[node + 1] - eb - 2;
  }
  for (uint32_t t = 0; t < ntri; t++) {
    V3 t0, t1, t2;
    if (is.has_height) {
      const uint8_t* tri = nav.htris + (size_t)(bt + t) * 3;
      t0 = ld3(nav.hverts, bv + tri[0]);
      t1 = ld3(nav.hverts, bv + tri[1]);
      t2 = ld3(nav.hverts, bv + tri[2]);
    } else {
      t0 = ld3(nav.verts_local, nav.edge_vertex[eb]);
      t1 = ld3(nav.verts_local, nav.edge_vertex[eb + t + 1]);
      t2 = ld3(nav.verts_local, nav.edge_vertex[eb + t + 2]);
    }
    V3 d0 = t1 - t0, d1 = t2 - t1, d2 = t0 - t2;
    if (perp_dot(xy(d0), xy(point) - xy(t0)) < EPS) continue;
    if (perp_dot(xy(d1), xy(point) - xy(t1)) < EPS) continue;
    if (perp_dot(xy(d2), xy(point) - xy(t2)) < EPS) continue;
    V3 normal = -normalize(cross(d0, d2));
    float height = fdiv(dot(normal, point - t0), normal.z);
    return {point.x, point.y, fsub(point.z, height)};
  }
  return point;  // the reference panics here; unreachable for consistent inputs
}

}  // namespace

// ---------------------------------------------------------------------------
__global__ void k_reset_slots(AgentArrays ag, AgentPersist persist, PathPool pool) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n) return;
  if (ag.flags[i] & LMB_AGENT_RESET) {
    uint32_t s = ag.slot[i];
    pool.slot_path_len[s] = 0;
    persist.state[s] = LMB_STATE_IDLE;
    persist.desired_move[3 * (size_t)s] = 0.0f;
    persist.desired_move[3 * (size_t)s + 1] = 0.0f;
    persist.desired_move[3 * (size_t)s + 2] = 0.0f;
    persist.reached_link_id[s] = LMB_NONE;
  }
}

// Phase A state writes (lib.rs:286-297) + phase B decision (lib.rs:345-424 minus the A* call).
// One WARP per agent: the two corridor searches of does_agent_need_repath are linear scans
// (path.rs:403-446) over corridors that reach tens of thousands of nodes on maze meshes, so the
// lanes scan 32 consecutive entries at a time (coalesced) with a ballot.
__global__ void __launch_bounds__(128)
k_repath_decide(DevNav nav, AgentArrays ag, AgentPersist persist, PathPool pool, RepathQueue queue,
                FollowScratch scratch) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (i >= ag.n) return;
  const uint32_t s = ag.slot[i];
  const uint32_t flags = ag.flags[i];
  const uint32_t plen = pool.slot_path_len[s];
  // decision: 0 none/keep, 1 follow, 2 repath  (lane-uniform)
  uint32_t new_state = LMB_NONE;
  bool clear_path = false, follow = false, repath = false;
  uint32_t ai = LMB_NONE, ti = LMB_NONE;
  const uint32_t agent_node = ag.agent_node[i], target_node = ag.target_node[i];
  // a path invalidated by the last navdata update is the marker entry (k_remap_paths)
  const bool invalid = plen && pool.entries[pool.slot_path_off[s]].x == LMB_NONE;
  if (flags & LMB_AGENT_PAUSED) {
    new_state = LMB_STATE_PAUSED;
    clear_path = invalid;  // lib.rs:350-358
  } else if (flags & LMB_AGENT_USING_ANIMATION_LINK) {
    new_state = LMB_STATE_USING_ANIMATION_LINK;
    clear_path = invalid;
  } else if (!(flags & LMB_AGENT_HAS_TARGET)) {  // does_agent_need_repath (agent.rs:425-475)
    if (plen) {  // ClearPathNoTarget; DoNothing leaves the state untouched
      new_state = LMB_STATE_IDLE;
      clear_path = true;
    }
  } else if (agent_node == LMB_NONE) {
    new_state = LMB_STATE_AGENT_NOT_ON_NAV_MESH;
    clear_path = true;
  } else if (target_node == LMB_NONE) {
    new_state = LMB_STATE_TARGET_NOT_ON_NAV_MESH;
    clear_path = true;
  } else {
    repath = true;
    if (plen && !invalid) {  // agent.rs:449-456
      const uint2* e = pool.entries + pool.slot_path_off[s];
      // find_index_of_node: first match from the front (path.rs:403-420)
      for (uint32_t base = 0; base < plen; base += 32) {
        uint32_t k = base + lane;
        uint32_t hit = __ballot_sync(0xffffffffu, k < plen && e[k].x == agent_node);
        if (hit) {
          ai = base + (uint32_t)__ffs(hit) - 1;
          break;
        }
      }
      if (ai != LMB_NONE) {
        // find_index_of_node_rev: first match from the back (path.rs:426-446)
        for (uint32_t end = plen; end > 0;) {
          uint32_t base = end >= 32 ? end - 32 : 0;
          uint32_t k = base + lane;
          uint32_t hit = __ballot_sync(0xffffffffu, k < end && e[k].x == target_node);
          if (hit) {
            ti = base + 31u - (uint32_t)__clz(hit);
            break;
          }
          end = base;
        }
        if (ti != LMB_NONE && !(ai > ti)) {
          follow = true;  // FollowPath
          repath = false;
        }
      }
    }
  }
  if (lane != 0) return;
  scratch.follow_start[i] = follow ? ai : LMB_NONE;
  scratch.follow_end[i] = follow ? ti : LMB_NONE;
  scratch.presult[i] = 0;
  scratch.query_of_agent[i] = LMB_NONE;
  scratch.waypoint_kind[i] = LMB_NONE;
  persist.reached_link_id[s] = LMB_NONE;  // current_animation_link = None (lib.rs:348)
  if (new_state != LMB_NONE) persist.state[s] = new_state;
  if (clear_path) pool.slot_path_len[s] = 0;
  if (!repath) return;
  // NeedsRepath
  pool.slot_path_len[s] = 0;
  uint32_t qi = atomicAdd(queue.count, 1u);
  queue.agent[qi] = i;
  queue.slot[qi] = s;
  queue.start_node[qi] = agent_node;
  queue.end_node[qi] = target_node;
  for (int k = 0; k < 3; k++) {
    queue.start_point[3 * (size_t)qi + k] = ag.agent_point[3 * (size_t)i + k];
    queue.end_point[3 * (size_t)qi + k] = ag.target_point[3 * (size_t)i + k];
  }
  scratch.query_of_agent[i] = qi;
}

// Post-A* bookkeeping (lib.rs:406-421) + phase C (lib.rs:426-523).
// ---- processing order of k_follow -----------------------------------------------------------------------
// The funnel scans the corridor from the agent to the first corner — on open ground to the corridor's end — so an
// agent's work is its remaining corridor length, anything from 0 to thousands of nodes, and a warp of agents in
// index order runs as long as its longest (config 3: 9 of 32 lanes active per instruction).  A counting sort by
// half-octave length bucket, longest first, gives warps of like lengths.
__device__ __forceinline__ uint32_t follow_bucket(uint32_t remaining) {
  if (remaining == 0) return 0;
  const uint32_t l = 31u - __clz(remaining);
  const uint32_t half = l ? (remaining >> (l - 1u)) & 1u : 0u;
  const uint32_t b = 1u + 2u * l + half;
  return b < FOLLOW_BUCKETS ? b : FOLLOW_BUCKETS - 1u;
}
__global__ void k_follow_order_count(AgentArrays ag, PathPool pool, AStarQueries q, FollowScratch scratch, uint8_t* key8,
                                     uint32_t* hist) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t key = FOLLOW_BUCKETS;  // no agent
  if (i < ag.n) {
    uint32_t remaining = 0;
    const uint32_t qi = scratch.query_of_agent[i];
    if (qi != LMB_NONE) {
      if (q.out_success[qi]) remaining = q.out_path_len[qi];
    } else if (pool.slot_path_len[ag.slot[i]]) {
      const uint32_t fs = scratch.follow_start[i], fe = scratch.follow_end[i];
      if (fs != LMB_NONE && fe != LMB_NONE && fe >= fs) remaining = fe - fs + 1u;
    }
    key = follow_bucket(remaining);
    key8[i] = (uint8_t)key;
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  if (key < FOLLOW_BUCKETS && (threadIdx.x & 31u) == (uint32_t)__ffs(peers) - 1u) atomicAdd(hist + key, __popc(peers));
}
// hist[b] -> first position of bucket b with the LONGEST bucket first, into hist[FOLLOW_BUCKETS + b] (the cursors)
__global__ void k_follow_order_offsets(uint32_t* hist) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int b = (int)FOLLOW_BUCKETS - 1; b >= 0; b--) {
      hist[FOLLOW_BUCKETS + b] = run;
      run += hist[b];
    }
  }
}
__global__ void k_follow_order_scatter(uint32_t n, const uint8_t* key8, uint32_t* hist, uint32_t* order) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t key = i < n ? key8[i] : FOLLOW_BUCKETS;
  const uint32_t peers = __match_any_sync(0xffffffffu, key);
  const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
  uint32_t base = 0;
  if (key < FOLLOW_BUCKETS && lane == leader) base = atomicAdd(hist + FOLLOW_BUCKETS + key, __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (key < FOLLOW_BUCKETS) order[base + __popc(peers & ((1u << lane) - 1u))] = i;
}

__global__ void __launch_bounds__(128)
k_follow(DevNav nav, AgentArrays ag, AgentPersist persist, PathPool pool, AStarQueries q,
         FollowScratch scratch) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ag.n) return;
  const uint32_t i = scratch.order ? scratch.order[t] : t;
  const uint32_t s = ag.slot[i];
  uint32_t follow_start = scratch.follow_start[i], follow_end = scratch.follow_end[i];
  const uint32_t qi = scratch.query_of_agent[i];
  if (qi != LMB_NONE) {
    uint32_t ok = q.out_success[qi];
    scratch.presult[i] = 0x80000000u | (ok ? 0x40000000u : 0u) | (q.out_explored[qi] & 0x3fffffffu);
    if (!ok) {
      persist.state[s] = LMB_STATE_NO_PATH;
    } else {
      follow_start = 0;                        // PathIndex::from_corridor_index(0, 0)
      follow_end = q.out_path_len[qi] - 1;     // new_path.last_index()
      float* ends = persist.path_endpoints + 6 * (size_t)s;  // Path { start_point, end_point }
      for (int k = 0; k < 3; k++) {
        ends[k] = q.start_point[3 * (size_t)qi + k];
        ends[3 + k] = q.end_point[3 * (size_t)qi + k];
      }
    }
  }
  const uint32_t plen = pool.slot_path_len[s];
  float* dm = persist.desired_move + 3 * (size_t)s;
  if (!plen) {
    dm[0] = dm[1] = dm[2] = 0.0f;
    return;
  }
  const uint32_t agent_node = ag.agent_node[i];
  if (agent_node == LMB_NONE || (ag.flags[i] & (LMB_AGENT_PAUSED | LMB_AGENT_USING_ANIMATION_LINK)))
    return;  // paused with a path keeps its old desired move (lib.rs:435-441)
  PathView path{pool.entries + pool.slot_path_off[s], plen};
  const V3 agent_point{ag.agent_point[3 * (size_t)i], ag.agent_point[3 * (size_t)i + 1], ag.agent_point[3 * (size_t)i + 2]};
  const V3 target_point{ag.target_point[3 * (size_t)i], ag.target_point[3 * (size_t)i + 1], ag.target_point[3 * (size_t)i + 2]};
  Step next;
  uint32_t next_index =
      find_next_point_in_straight_path(nav, path, follow_start, agent_point, follow_end, target_point, next);
  {
    scratch.waypoint_kind[i] = next.is_link ? 1u : 0u;
    scratch.waypoint_link[i] = next.is_link ? next.link_id : LMB_NONE;
    float* w = scratch.waypoint_points + 6 * (size_t)i;
    w[0] = next.point.x; w[1] = next.point.y; w[2] = next.point.z;
    const V3 e = next.is_link ? next.end_point : next.point;
    w[3] = e.x; w[4] = e.y; w[5] = e.z;
  }
  const float radius = ag.radius[i];
  const uint32_t cond = ag.reach_condition ? ag.reach_condition[i] : LMB_REACH_DISTANCE;
  const float rd = reach_distance_or_radius(ag.reach_distance ? ag.reach_distance[i] : -1.0f, radius);
  if (has_reached_target(nav, path, cond, rd, agent_point, next_index, next, follow_end, target_point)) {
    dm[0] = dm[1] = dm[2] = 0.0f;
    persist.state[s] = LMB_STATE_REACHED_TARGET;
    return;
  }
  V3 waypoint;
  if (!next.is_link) {
    persist.state[s] = LMB_STATE_MOVING;
    waypoint = next.point;
  } else {
    // refine the link's end points onto the nav mesh (lib.rs:483-498)
    const DevIsland& is0 = nav.islands[nav.poly_island[next.start_node]];
    V3 start_point = island_apply(is0, sample_point_on_node(nav, island_apply_inverse(is0, next.point), next.start_node));
    const DevIsland& is1 = nav.islands[nav.poly_island[next.end_node]];
    V3 end_point = island_apply(is1, sample_point_on_node(nav, island_apply_inverse(is1, next.end_point), next.end_node));
    float d = distance(agent_point, start_point);
    float alr = ag.anim_link_reach_distance ? ag.anim_link_reach_distance[i] : -1.0f;
    float link_reach = alr >= 0.0f ? alr : rd;  // agent.rs:312-323
    if (d <= link_reach) {
      persist.state[s] = LMB_STATE_REACHED_ANIMATION_LINK;
      persist.reached_link_id[s] = next.link_id;
      float* a = persist.reached_link_start + 3 * (size_t)s;
      float* b = persist.reached_link_end + 3 * (size_t)s;
      a[0] = start_point.x; a[1] = start_point.y; a[2] = start_point.z;
      b[0] = end_point.x; b[1] = end_point.y; b[2] = end_point.z;
    } else {
      persist.state[s] = LMB_STATE_MOVING;
    }
    waypoint = start_point;
  }
  const V3 position{ag.position[3 * (size_t)i], ag.position[3 * (size_t)i + 1], ag.position[3 * (size_t)i + 2]};
  V2 move = normalize_or_zero(xy(waypoint - position)) * ag.desired_speed[i];
  dm[0] = move.x;
  dm[1] = move.y;
  dm[2] = 0.0f;
}

// Batched `Archipelago::find_path` straight path (query.rs:218-272): repeated funnel calls from
// the start of the corridor to its end.  One thread per query; steps are written to
// [step_begin[i], step_begin[i] + step_count[i]).
__global__ void __launch_bounds__(128)
k_straight_paths(DevNav nav, PathPool pool, AStarQueries q, const uint32_t* __restrict__ step_begin,
                 uint32_t* __restrict__ step_count, uint32_t* __restrict__ out_kind,
                 float* __restrict__ out_points, uint32_t* __restrict__ out_link) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n) return;
  step_count[i] = 0;
  if (!q.out_success[i]) return;
  const uint32_t cap = step_begin[i + 1] - step_begin[i];
  uint32_t base = step_begin[i], count = 0;
  auto emit = [&](uint32_t kind, V3 a, V3 b, uint32_t link) {
    if (count < cap) {
      uint32_t k = base + count;
      out_kind[k] = kind;
      out_points[6 * (size_t)k + 0] = a.x;
      out_points[6 * (size_t)k + 1] = a.y;
      out_points[6 * (size_t)k + 2] = a.z;
      out_points[6 * (size_t)k + 3] = b.x;
      out_points[6 * (size_t)k + 4] = b.y;
      out_points[6 * (size_t)k + 5] = b.z;
      out_link[k] = link;
    }
    count++;
  };
  PathView path{pool.entries + q.out_path_off[i], q.out_path_len[i]};
  V3 current_point{q.start_point[3 * (size_t)i], q.start_point[3 * (size_t)i + 1], q.start_point[3 * (size_t)i + 2]};
  const V3 last_point{q.end_point[3 * (size_t)i], q.end_point[3 * (size_t)i + 1], q.end_point[3 * (size_t)i + 2]};
  uint32_t current_index = 0;
  const uint32_t last_index = path.len - 1;
  emit(0u, current_point, current_point, LMB_NONE);
  if (current_index == last_index) {
    emit(0u, last_point, last_point, LMB_NONE);
    step_count[i] = count;
    return;
  }
  bool last_was_link = false;
  while ((current_index != last_index || last_was_link) && count < cap) {
    Step next;
    current_index = find_next_point_in_straight_path(nav, path, current_index, current_point, last_index,
                                                     last_point, next);
    if (!next.is_link) {
      current_point = next.point;
      emit(0u, next.point, next.point, LMB_NONE);
      last_was_link = false;
    } else {
      current_point = next.end_point;
      emit(1u, next.point, next.end_point, next.link_id);
      last_was_link = true;
    }
  }
  step_count[i] = count;
}

void launch_straight_paths(const DevNav& nav, const PathPool& pool, const AStarQueries& q,
                           const uint32_t* step_begin, uint32_t* step_count, uint32_t* out_kind,
                           float* out_points, uint32_t* out_link, cudaStream_t stream) {
  if (q.n)
    k_straight_paths<<<(q.n + 127) / 128, 128, 0, stream>>>(nav, pool, q, step_begin, step_count, out_kind,
                                                            out_points, out_link);
}

__global__ void k_gather_outputs(AgentArrays ag, AgentPersist persist, float* out_velocity,
                                 uint32_t* out_state, uint32_t* out_link_id, float* out_link_start,
                                 float* out_link_end) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n) return;
  uint32_t s = ag.slot[i];
  for (int k = 0; k < 3; k++) out_velocity[3 * (size_t)i + k] = persist.desired_move[3 * (size_t)s + k];
  out_state[i] = persist.state[s];
  uint32_t lid = persist.reached_link_id[s];
  out_link_id[i] = lid;
  for (int k = 0; k < 3; k++) {
    out_link_start[3 * (size_t)i + k] = lid != LMB_NONE ? persist.reached_link_start[3 * (size_t)s + k] : 0.0f;
    out_link_end[3 * (size_t)i + k] = lid != LMB_NONE ? persist.reached_link_end[3 * (size_t)s + k] : 0.0f;
  }
}

__global__ void k_drop_all_paths(PathPool pool, uint32_t max_agents) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < max_agents) pool.slot_path_len[i] = 0;
  if (i == 0) *pool.cursor = 0ull;
}

// Carries stored paths over a navdata replacement (`Path::is_valid`, path.rs:381-400, against the
// sets `NavigationData::update` returns).  One warp per slot.  A path none of whose nodes lies on
// an invalidated island and none of whose off-mesh links was dropped gets its node ids and link
// indices rewritten to the new flattening; any other path becomes the one-entry marker
// (LMB_NONE, LMB_NONE), which k_repath_decide treats as "has a path, path invalid".
__global__ void __launch_bounds__(128)
k_remap_paths(PathPool pool, uint32_t max_agents, const uint32_t* __restrict__ old_begin, uint32_t n_old_islands,
              const uint32_t* __restrict__ new_begin_of_old, const uint32_t* __restrict__ link_map,
              uint32_t n_old_links) {
  const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (slot >= max_agents) return;
  const uint32_t len = pool.slot_path_len[slot];
  if (!len) return;
  uint2* e = pool.entries + pool.slot_path_off[slot];
  bool bad = false;
  for (uint32_t k = lane; k < len; k += 32) {
    uint2 v = e[k];
    if (v.x == LMB_NONE) {
      bad = true;
      break;
    }
    uint32_t lo = 0, hi = n_old_islands;  // last island whose first node id is <= v.x
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (old_begin[mid] <= v.x) lo = mid; else hi = mid;
    }
    const uint32_t nb = new_begin_of_old[lo];
    if (nb == LMB_NONE || v.x >= old_begin[lo + 1]) {
      bad = true;
      break;
    }
    v.x = nb + (v.x - old_begin[lo]);
    if (v.y != LMB_NONE && (v.y & LMB_PORTAL_LINK)) {
      const uint32_t l = v.y & ~LMB_PORTAL_LINK;
      const uint32_t nl = l < n_old_links ? link_map[l] : LMB_NONE;
      if (nl == LMB_NONE) {
        bad = true;
        break;
      }
      v.y = LMB_PORTAL_LINK | nl;
    }
    e[k] = v;
  }
  bad = __any_sync(0xffffffffu, bad);
  if (bad && lane == 0) {
    e[0] = make_uint2(LMB_NONE, LMB_NONE);
    pool.slot_path_len[slot] = 1;
  }
}

// ---- single-block exclusive scan ------------------------------------------------------
__global__ void __launch_bounds__(1024) k_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= (uint32_t)o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    uint32_t carry = carry_s;
    uint32_t prefix = carry + (warp ? warp_sums[warp - 1] : 0u) + x - v;
    if (i < n) out[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
}

// Multi-block form (reduce, scan the tile sums, scan the tiles): the avoidance grid has up to 8 M cells and
// one block walking them 1024 at a time was a visible part of a steady frame.
constexpr uint32_t SCAN_TILE = 4096;  // 1024 threads x 4 consecutive elements

__global__ void __launch_bounds__(1024) k_scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* sums) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t i0 = (size_t)blockIdx.x * SCAN_TILE + 4u * threadIdx.x;
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (i0 + k < n) v += in[i0 + k];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane];
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) sums[blockIdx.x] = w;
  }
}

__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* in, uint32_t* out, uint32_t n,
                                                     const uint32_t* __restrict__ tile_offset) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t i0 = (size_t)blockIdx.x * SCAN_TILE + 4u * threadIdx.x;
  uint32_t v[4], t = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    v[k] = i0 + k < n ? in[i0 + k] : 0u;
    t += v[k];
  }
  uint32_t x = t;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= (uint32_t)o) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  uint32_t prefix = tile_offset[blockIdx.x] + (warp ? warp_sums[warp - 1] : 0u) + x - t;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (i0 + k < n) out[i0 + k] = prefix;
    prefix += v[k];
  }
}

uint32_t exclusive_scan_tmp_words(uint32_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t stream) {
  if (n == 0) return;
  if (n <= SCAN_TILE || !tmp) {
    k_exclusive_scan<<<1, 1024, 0, stream>>>(in, out, n);
    return;
  }
  const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_tile_sums<<<tiles, 1024, 0, stream>>>(in, n, tmp);
  k_exclusive_scan<<<1, 1024, 0, stream>>>(tmp, tmp, tiles);
  k_scan_tiles<<<tiles, 1024, 0, stream>>>(in, out, n, tmp);
}

// What a caller's movement system does between two updates (README example; SURVEY.md §8d common inputs:
// velocity = previous desired velocity, p += v dt), on the device, so that a resident loop needs no host round
// trip: position += desired_velocity * dt, velocity = desired_velocity, LMB_AGENT_RESET cleared.
__global__ void k_integrate(uint32_t n, float dt, const float* __restrict__ out_velocity, float* position,
                            float* velocity, uint32_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < 3; k++) {
    const float v = out_velocity[3 * (size_t)i + k];
    position[3 * (size_t)i + k] = fadd(position[3 * (size_t)i + k], fmul(v, dt));
    velocity[3 * (size_t)i + k] = v;
  }
  flags[i] &= ~LMB_AGENT_RESET;
}

void launch_integrate(uint32_t n, float dt, const float* out_velocity, float* position, float* velocity,
                      uint32_t* flags, cudaStream_t stream) {
  if (n) k_integrate<<<(n + 255) / 256, 256, 0, stream>>>(n, dt, out_velocity, position, velocity, flags);
}

// ---- path pool compaction -----------------------------------------------------------
__global__ void k_pool_compact_copy(PathPool src, uint2* dst, const uint32_t* new_off, uint32_t max_agents) {
  // one warp per slot
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= max_agents) return;
  uint32_t len = src.slot_path_len[warp];
  if (!len) return;
  uint32_t so = src.slot_path_off[warp], d = new_off[warp];
  for (uint32_t k = lane; k < len; k += 32) dst[d + k] = src.entries[so + k];
  __syncwarp();
  if (lane == 0) src.slot_path_off[warp] = d;
}
__global__ void k_pool_set_cursor(PathPool pool, const uint32_t* new_off, uint32_t max_agents) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *pool.cursor = (unsigned long long)new_off[max_agents - 1] + pool.slot_path_len[max_agents - 1];
}

// step 1: new offsets (exclusive scan of the lengths) and the live total into *cursor
void launch_pool_scan(const PathPool& src, uint32_t max_agents, uint32_t* scan_tmp, uint32_t* tile_tmp,
                      cudaStream_t stream) {
  launch_exclusive_scan(src.slot_path_len, scan_tmp, max_agents, tile_tmp, stream);
  k_pool_set_cursor<<<1, 32, 0, stream>>>(src, scan_tmp, max_agents);
}
// step 2: copy every live path to its new offset in dst and update slot_path_off
void launch_pool_copy(const PathPool& src, uint2* dst_entries, uint32_t max_agents, const uint32_t* scan_tmp,
                      cudaStream_t stream) {
  uint32_t threads = 256, warps_per_block = threads / 32;
  k_pool_compact_copy<<<(max_agents + warps_per_block - 1) / warps_per_block, threads, 0, stream>>>(
      src, dst_entries, scan_tmp, max_agents);
}

// ---- launchers -------------------------------------------------------------------------
static inline uint32_t blocks_for(uint32_t n, uint32_t b) { return (n + b - 1) / b; }

void launch_reset_slots(const AgentArrays& ag, const AgentPersist& persist, const PathPool& pool,
                        cudaStream_t stream) {
  if (ag.n) k_reset_slots<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, persist, pool);
}
void launch_repath_decide(const DevNav& nav, const AgentArrays& ag, const AgentPersist& persist,
                          const PathPool& pool, const RepathQueue& queue, const FollowScratch& scratch,
                          cudaStream_t stream) {
  if (ag.n) k_repath_decide<<<blocks_for(ag.n, 4), 128, 0, stream>>>(nav, ag, persist, pool, queue, scratch);
}
void launch_follow(const DevNav& nav, const AgentArrays& ag, const AgentPersist& persist,
                   const PathPool& pool, const AStarQueries& q, const FollowScratch& scratch,
                   cudaStream_t stream) {
  if (ag.n) k_follow<<<blocks_for(ag.n, 128), 128, 0, stream>>>(nav, ag, persist, pool, q, scratch);
}
void launch_follow_order(const AgentArrays& ag, const PathPool& pool, const AStarQueries& q, const FollowScratch& scratch,
                         uint8_t* key8, uint32_t* hist, uint32_t* order, cudaStream_t stream) {
  if (!ag.n) return;
  cudaMemsetAsync(hist, 0, 2 * FOLLOW_BUCKETS * sizeof(uint32_t), stream);
  k_follow_order_count<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, pool, q, scratch, key8, hist);
  k_follow_order_offsets<<<1, 32, 0, stream>>>(hist);
  k_follow_order_scatter<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag.n, key8, hist, order);
}
void launch_gather_outputs(const AgentArrays& ag, const AgentPersist& persist, float* out_velocity,
                           uint32_t* out_state, uint32_t* out_link_id, float* out_link_start,
                           float* out_link_end, cudaStream_t stream) {
  if (ag.n)
    k_gather_outputs<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, persist, out_velocity, out_state,
                                                                out_link_id, out_link_start, out_link_end);
}
void launch_drop_all_paths(const PathPool& pool, uint32_t max_agents, cudaStream_t stream) {
  k_drop_all_paths<<<blocks_for(max_agents ? max_agents : 1, 256), 256, 0, stream>>>(pool, max_agents);
}
void launch_remap_paths(const PathPool& pool, uint32_t max_agents, const uint32_t* old_begin, uint32_t n_old_islands,
                        const uint32_t* new_begin_of_old, const uint32_t* link_map, uint32_t n_old_links,
                        cudaStream_t stream) {
  if (max_agents && n_old_islands)
    k_remap_paths<<<blocks_for(max_agents, 4), 128, 0, stream>>>(pool, max_agents, old_begin, n_old_islands,
                                                                 new_begin_of_old, link_map, n_old_links);
}

}  // namespace lmb
