// TEST INFRASTRUCTURE — internal types of the CPU oracle (see oracle/README.md).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <optional>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../include/landmass_b200.h"
#include "orc_math.h"

namespace orc {

struct Island {
  Xform xf;
  V3 bmin, bmax;
  bool bounds_empty;
  uint32_t poly_begin, poly_end, vert_begin, vert_end;
  bool has_height;
};

// PermittedAnimationLinks (agent.rs:170-187)
struct Permitted {
  bool all = true;
  std::unordered_set<uint32_t> kinds;
  bool is_permitted(uint32_t kind) const { return all || kinds.count(kind) != 0; }
};

struct ModifiedNode {  // nav_data.rs:119-132
  std::vector<std::pair<uint32_t, uint32_t>> new_boundary;
  std::vector<V2> new_vertices;
};

struct PossibleLink {
  uint32_t kind, dst;
};

struct NavData {
  uint32_t n_islands = 0, n_polys = 0, n_links = 0;
  std::vector<Island> islands;
  std::vector<V3> vertices;
  std::vector<uint32_t> poly_edge_begin, edge_vertex, edge_neighbor, edge_reverse, poly_type,
      poly_region, poly_island;
  std::vector<float> poly_bounds, poly_center;
  std::vector<uint32_t> hpoly;
  std::vector<V3> hverts;
  std::vector<uint8_t> htris;
  std::map<uint32_t, float> type_cost;
  std::vector<uint32_t> link_src_node, link_dst_node, link_dst_type, link_kind, link_reverse,
      link_anim_kind, link_anim_id, node_link_begin, node_links;
  std::vector<float> link_cost;
  std::vector<std::pair<V3, V3>> link_portal, link_dst_portal;
  std::unordered_map<uint32_t, ModifiedNode> modified;
  std::vector<uint32_t> node_region_number, region_root;
  std::unordered_map<uint32_t, std::vector<PossibleLink>> possible_links;

  void load(const lmb_navdata_desc& d);
  void edge_indices(uint32_t node, uint32_t edge, uint32_t& left, uint32_t& right) const;
  bool sample_point_island(uint32_t island, V3 point, const lmb_options& o, V3& out_point,
                           uint32_t& out_node) const;
  bool sample_point(V3 point, const lmb_options& o, V3& out_point, uint32_t& out_node) const;
  bool sample_point_on_node(V3 point, uint32_t node, V3& out) const;
  bool are_nodes_connected(uint32_t n1, uint32_t n2, const Permitted& permitted) const;
  V3 world_vertex(uint32_t node, uint32_t global_vertex) const {
    return islands[poly_island[node]].xf.apply(vertices[global_vertex]);
  }
};

// --- A* -------------------------------------------------------------------
struct HeapEntry {
  float cost, estimate;
  uint32_t index;
};
struct RustHeap {  // std::collections::BinaryHeap<Reverse<NodeRef>>
  std::vector<HeapEntry> data;
  void sift_up(size_t start, size_t pos);
  void push(const HeapEntry& e);
  bool pop(HeapEntry& out);
};
struct Successor {
  float cost;
  int64_t action;
  uint64_t state;
};
struct AStarResult {
  bool found = false;
  uint32_t explored_nodes = 0;
  std::vector<int64_t> actions;
};

// --- Path -------------------------------------------------------------------
struct Path;
struct PathIndex {  // path.rs:134-177
  uint32_t segment_index, portal_index;
  bool operator==(const PathIndex& o) const {
    return segment_index == o.segment_index && portal_index == o.portal_index;
  }
  bool operator<=(const PathIndex& o) const {
    return segment_index < o.segment_index ||
           (segment_index == o.segment_index && portal_index <= o.portal_index);
  }
  bool operator>(const PathIndex& o) const { return !(*this <= o); }
  PathIndex next(const Path& path) const;
};
struct IslandSegment {
  uint32_t island;
  std::vector<uint32_t> corridor;  // GLOBAL node ids
  std::vector<uint32_t> portal_edge_index;
};
struct OffMeshLinkSegment {
  uint32_t starting_node, end_node, link;
};
struct StraightPathStep {  // path.rs:181-199
  bool is_link;
  V3 point;  // Waypoint point / AnimationLink start_point
  V3 end_point;
  uint32_t link_id, start_node, end_node;
};
struct Path {
  std::vector<IslandSegment> island_segments;
  std::vector<OffMeshLinkSegment> off_mesh_link_segments;
  V3 start_point, end_point;
  PathIndex last_index() const;
  // path.rs:381-400 against the sets of an lmb_nav_remap (indices of the flattening the path is in)
  bool is_valid(const uint32_t* island_map, const uint32_t* link_map) const;
  std::optional<PathIndex> find_index_of_node(const NavData& nav, uint32_t node) const;
  std::optional<PathIndex> find_index_of_node_rev(const NavData& nav, uint32_t node) const;
  std::pair<PathIndex, StraightPathStep> find_next_point_in_straight_path(
      const NavData& nav, PathIndex start_index, V3 start_point, PathIndex end_index,
      V3 end_point) const;
  void flatten(std::vector<uint32_t>& nodes, std::vector<uint32_t>& portals) const;
  static Path from_flat(const NavData& nav, uint32_t len, const uint32_t* nodes,
                        const uint32_t* portals);
  uint32_t flat_index(PathIndex idx) const;
  PathIndex from_flat_index(uint32_t flat) const;
};
struct PathResult {
  uint32_t explored_nodes = 0;
  std::optional<Path> path;
};
PathResult find_path(const NavData& nav, uint32_t start_node, V3 start_point, uint32_t end_node,
                     V3 end_point, const std::unordered_map<uint32_t, float>& overrides,
                     const Permitted& permitted);

// --- agents -----------------------------------------------------------------
struct AgentIn {
  uint32_t slot, flags;
  V3 position, velocity, target{0, 0, 0};
  float radius, desired_speed, max_speed;
  uint32_t reach_condition;
  float reach_distance, anim_link_reach_distance;
  std::unordered_map<uint32_t, float> overrides;
  Permitted permitted;
};
struct CharacterIn {
  V3 position, velocity, point;
  float radius;
  bool valid;
};
struct Sampled {
  bool valid = false;
  V3 point{0, 0, 0};
  uint32_t node = LMB_NONE;
};
struct AgentPersist {  // private Agent fields (agent.rs:93-104)
  std::optional<Path> path;
  // `!path.is_valid(invalidated_off_mesh_links, invalidated_islands)` of the navdata update that
  // precedes the next step (the sets only exist for that one frame: lib.rs:274-281)
  bool path_invalid = false;
  V3 desired_move{0, 0, 0};
  uint32_t state = LMB_STATE_IDLE;
  uint32_t reached_link_id = LMB_NONE;
  V3 reached_link_start{0, 0, 0}, reached_link_end{0, 0, 0};
};

// avoidance_oracle.cpp
struct Obstacle {
  bool closed;
  std::vector<V2> vertices;
};
std::vector<Obstacle> nav_mesh_borders_to_dodgy_obstacles(const NavData& nav, V3 agent_point,
                                                          uint32_t agent_node, float distance_limit);
struct AvoidRec {  // 32 bytes, same layout as the device halo record
  V3 p;
  V2 v;
  float radius, responsibility;
  uint32_t valid;
};
void build_avoid_records(const std::vector<AgentIn>& agents, const std::vector<AgentPersist*>& persist,
                         const std::vector<Sampled>& agent_node, const lmb_options& opt, AvoidRec* out);
// debug-avoidance export (debug.rs:415-503): lines as {point.xy, direction.xy}
struct AvoidDebugData {
  bool ran = false, fell_back = false;
  std::vector<std::array<float, 4>> original, fallback;
};
void apply_avoidance_to_agents(const NavData& nav, const std::vector<AvoidRec>& records, uint32_t first,
                               const std::vector<AgentIn>& agents,
                               const std::vector<AgentPersist*>& persist,
                               const std::vector<Sampled>& agent_node,
                               const std::vector<CharacterIn>& characters, const lmb_options& opt,
                               float delta_time, std::vector<AvoidDebugData>* debug = nullptr);

struct Ctx {
  uint32_t flags = 0;
  NavData nav;
  bool has_nav = false;
  std::unordered_map<uint32_t, AgentPersist> agent_state;
  std::vector<lmb_pathing_result> pathing_results;
  // frame state kept between step_begin and step_end (multi-rank form)
  std::vector<AgentIn> f_agents;
  std::vector<Sampled> f_agent_node;
  std::vector<CharacterIn> f_chars;
  std::vector<AvoidRec> halo;
  // the frame's next step of the straight path per agent (debug.rs:386-412); kind LMB_NONE = none
  struct Waypoint {
    uint32_t kind = LMB_NONE, link = LMB_NONE;
    V3 point{0, 0, 0}, end_point{0, 0, 0};
  };
  std::vector<Waypoint> f_waypoints;
  std::vector<AvoidDebugData> f_avoid_debug;  // per agent of the last step
  uint32_t halo_first = 0;
  void step_begin(const lmb_agents_in& in, const lmb_characters_in* chars, const lmb_options& opt, float dt,
                  uint32_t n_total, uint32_t first);
  void step_end(const lmb_options& opt, float dt, const lmb_agents_out& out);
  void step(const lmb_agents_in& in, const lmb_characters_in* chars, const lmb_options& opt, float dt,
            const lmb_agents_out& out) {
    step_begin(in, chars, opt, dt, in.n, 0);
    step_end(opt, dt, out);
  }
};

}  // namespace orc

struct orc_ctx {
  orc::Ctx c;
};
