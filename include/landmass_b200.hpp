// landmass_b200.hpp — C++17 host mirror of the reference's per-frame surface, over the C ABI.
//
// The reference is a Rust crate; this image has no Rust toolchain, so the host side that sits above
// include/landmass_b200.h exists twice: the complete mirror in Python (landmass_b200/*.py: validation,
// island linking, animation links, queries, debug drawing — what the parity tests drive) and this
// header, which covers the part of the API a game loop touches every frame, in the compiled language
// a maintainer would port to Rust line by line:
//
//   Archipelago<..>::{add,remove,get}_{agent,character}, update(dt), get_pathing_results,  lib.rs:96-268, 270-534
//   sample_point, find_path (SampledPoint, PathStep)                                       query.rs:19-127, 185-273
//   Agent (pub fields + state accessors), AgentState, TargetReachedCondition,              agent.rs:24-310
//   PermittedAnimationLinks, ReachedAnimationLink
//   Character                                                                              character.rs:14-31
//   ArchipelagoOptions::from_agent_radius                                                  lib.rs:63-93, coords.rs:148-158
//   DenseSlotMap (iteration = dense order, swap-remove, versioned keys)                     slotmap 1.0.7
//   NavigationMesh::validate -> ValidNavigationMesh, ValidationError, Transform              nav_mesh.rs:19-35, 131-407
//
// What `update` does here is what `Archipelago::update` keeps on the host: pack the
// agents in DenseSlotMap order, one lmb_step, scatter desired velocity / state / reached link back,
// translate `PathingResult.agent` from an index to the AgentId.  Navigation data enters as the
// flattened `lmb_navdata_desc`: `NavigationMesh::validate` (height meshes included) and a single-island
// `NavigationData` are mirrored below (SingleIslandNavigationData), worlds of several linked islands in
// landmass_b200_linking.hpp; only worlds with ANIMATION LINKS bring the descriptor the Python mirror (or
// the caller's port of nav_data.rs:516-825) flattens.
//
// Coordinates are landmass-internal (XYZ, z up): `CoordinateSystem::to_landmass` stays with the caller.
// The entry points are reached through a `Backend` policy class so that the same code runs on the
// CUDA library (`CudaBackend`, the only product backend — there is no CPU fallback) and, in tests
// only, on the CPU oracle, whose step functions have the same signatures.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "landmass_b200.h"

namespace landmass_b200 {

using Vec3 = std::array<float, 3>;

// agent.rs:24-47 (same order as the Rust enum and as lmb_agent_state)
enum class AgentState : uint32_t {
  Idle = LMB_STATE_IDLE,
  ReachedTarget = LMB_STATE_REACHED_TARGET,
  ReachedAnimationLink = LMB_STATE_REACHED_ANIMATION_LINK,
  UsingAnimationLink = LMB_STATE_USING_ANIMATION_LINK,
  Moving = LMB_STATE_MOVING,
  AgentNotOnNavMesh = LMB_STATE_AGENT_NOT_ON_NAV_MESH,
  TargetNotOnNavMesh = LMB_STATE_TARGET_NOT_ON_NAV_MESH,
  NoPath = LMB_STATE_NO_PATH,
  Paused = LMB_STATE_PAUSED,
};

// agent.rs:136-160
struct TargetReachedCondition {
  uint32_t kind = LMB_REACH_DISTANCE;
  std::optional<float> distance;  // None = the agent's radius
  static TargetReachedCondition Distance(std::optional<float> d = std::nullopt) { return {LMB_REACH_DISTANCE, d}; }
  static TargetReachedCondition VisibleAtDistance(std::optional<float> d = std::nullopt) {
    return {LMB_REACH_VISIBLE_AT_DISTANCE, d};
  }
  static TargetReachedCondition StraightPathDistance(std::optional<float> d = std::nullopt) {
    return {LMB_REACH_STRAIGHT_PATH_DISTANCE, d};
  }
};

// agent.rs:170-187
struct PermittedAnimationLinks {
  bool all = true;
  std::vector<uint32_t> kinds;  // used when !all; kept ascending when packed
};

// agent.rs:112-119
struct ReachedAnimationLink {
  uint32_t link_id;
  Vec3 start_point, end_point;
};

struct NotReachedAnimationLinkError : std::runtime_error {  // agent.rs:477-487
  NotReachedAnimationLinkError() : std::runtime_error("the agent has not reached an animation link") {}
};
struct NotUsingAnimationLinkError : std::runtime_error {
  NotUsingAnimationLinkError() : std::runtime_error("the agent is not using an animation link") {}
};

// agent.rs:50-310
class Agent {
 public:
  Vec3 position{}, velocity{};
  float radius = 0.0f, desired_speed = 0.0f, max_speed = 0.0f;
  std::optional<Vec3> current_target;
  TargetReachedCondition target_reached_condition = TargetReachedCondition::Distance();
  std::optional<float> animation_link_reached_distance;
  PermittedAnimationLinks permitted_animation_links;
  bool paused = false;
  bool keep_avoidance_data = false;  // feature `debug-avoidance`; read back with lmb_agent_avoidance_constraints

  static Agent create(Vec3 position, Vec3 velocity, float radius, float desired_speed, float max_speed) {
    Agent a;
    a.position = position;
    a.velocity = velocity;
    a.radius = radius;
    a.desired_speed = desired_speed;
    a.max_speed = max_speed;
    return a;
  }
  // agent.rs:232-262: costs must be positive; returns whether the override was taken / existed
  bool override_type_index_cost(uint32_t type_index, float cost) {
    if (!(cost > 0.0f)) return false;
    for (auto& kv : overrides_)
      if (kv.first == type_index) {
        kv.second = cost;
        return true;
      }
    overrides_.emplace_back(type_index, cost);
    return true;
  }
  bool remove_overridden_type_index_cost(uint32_t type_index) {
    for (size_t i = 0; i < overrides_.size(); i++)
      if (overrides_[i].first == type_index) {
        overrides_.erase(overrides_.begin() + (long)i);
        return true;
      }
    return false;
  }
  const std::vector<std::pair<uint32_t, float>>& get_type_index_cost_overrides() const { return overrides_; }
  const Vec3& get_desired_velocity() const { return current_desired_move_; }
  AgentState state() const { return state_; }
  const std::optional<ReachedAnimationLink>& reached_animation_link() const { return current_animation_link_; }
  void start_animation_link() {  // agent.rs:277-286
    if (!current_animation_link_) throw NotReachedAnimationLinkError();
    using_animation_link_ = true;
  }
  void end_animation_link() {  // agent.rs:290-299
    if (!using_animation_link_) throw NotUsingAnimationLinkError();
    using_animation_link_ = false;
  }
  bool is_using_animation_link() const { return using_animation_link_; }
  // feature `debug-avoidance` (agent.rs:105-108, debug.rs:415-503): the constraints of the last update that ran
  // avoidance for this agent while keep_avoidance_data was set; lines are {point.xy, direction.xy}
  // (ConstraintLine.normal = (-direction.y, direction.x))
  struct AvoidanceData {
    bool fell_back = false;  // DebugData::Fallback (else Satisfied: `fallback` is empty)
    std::vector<std::array<float, 4>> original, fallback;
  };
  const AvoidanceData* avoidance_data() const { return has_avoidance_data_ ? &avoidance_data_ : nullptr; }

 private:
  template <class B>
  friend class BasicArchipelago;
  std::vector<std::pair<uint32_t, float>> overrides_;
  Vec3 current_desired_move_{};
  AgentState state_ = AgentState::Idle;
  std::optional<ReachedAnimationLink> current_animation_link_;
  bool using_animation_link_ = false;
  AvoidanceData avoidance_data_;
  bool has_avoidance_data_ = false;
};

// character.rs:14-31
struct Character {
  Vec3 position{}, velocity{};
  float radius = 0.0f;
};

// lib.rs:63-93 with PointSampleDistance3d::from_agent_radius (coords.rs:148-158) already folded into
// CorePointSampleDistance (coords.rs:209-226)
struct ArchipelagoOptions {
  lmb_options abi{};
  static ArchipelagoOptions from_agent_radius(float radius) {
    ArchipelagoOptions o;
    o.abi.horizontal_distance = 0.2f * radius;
    o.abi.distance_above = 0.5f * radius;
    o.abi.distance_below = radius;
    o.abi.vertical_preference_ratio = 2.0f;
    o.abi.neighbourhood = 10.0f * radius;
    o.abi.avoidance_time_horizon = 0.5f;
    o.abi.obstacle_avoidance_time_horizon = 0.25f;
    o.abi.reached_destination_avoidance_responsibility = 0.1f;
    return o;
  }
};

// slotmap::KeyData: index + version (occupied versions are odd)
struct Key {
  uint32_t idx = 0, version = 0;
  bool operator==(const Key& o) const { return idx == o.idx && version == o.version; }
  bool operator!=(const Key& o) const { return !(*this == o); }
};
using AgentId = Key;
using CharacterId = Key;

// slotmap::DenseSlotMap: iteration order = dense storage order, removal swaps the last element in.
// This order is the agent order of the step, of `pathing_results` and of avoidance's tie-breaks.
template <class T>
class DenseSlotMap {
 public:
  DenseSlotMap() { slots_.push_back({0u, 0u}); }  // slot 0 is slotmap's sentinel
  Key insert(T value) {
    const uint32_t dense = (uint32_t)keys_.size();
    const uint32_t idx = free_head_;
    if (idx < slots_.size()) {
      free_head_ = slots_[idx].dense_or_next;
      slots_[idx].version |= 1u;
      slots_[idx].dense_or_next = dense;
    } else {
      slots_.push_back({1u, dense});
      free_head_ = (uint32_t)slots_.size();
    }
    const Key key{idx, slots_[idx].version};
    keys_.push_back(key);
    values_.push_back(std::move(value));
    return key;
  }
  std::optional<T> remove(Key key) {
    const int64_t d = dense_index(key);
    if (d < 0) return std::nullopt;
    slots_[key.idx].version += 1u;  // even = vacant
    slots_[key.idx].dense_or_next = free_head_;
    free_head_ = key.idx;
    T value = std::move(values_[(size_t)d]);
    const Key last = keys_.back();
    keys_[(size_t)d] = last;
    values_[(size_t)d] = std::move(values_.back());
    keys_.pop_back();
    values_.pop_back();
    if ((size_t)d < keys_.size()) slots_[last.idx].dense_or_next = (uint32_t)d;
    return value;
  }
  T* get(Key key) {
    const int64_t d = dense_index(key);
    return d < 0 ? nullptr : &values_[(size_t)d];
  }
  const T* get(Key key) const {
    const int64_t d = dense_index(key);
    return d < 0 ? nullptr : &values_[(size_t)d];
  }
  int64_t dense_index(Key key) const {
    if (key.idx >= slots_.size() || slots_[key.idx].version != key.version) return -1;
    return (int64_t)slots_[key.idx].dense_or_next;
  }
  size_t size() const { return keys_.size(); }
  const std::vector<Key>& keys() const { return keys_; }
  std::vector<T>& values() { return values_; }
  const std::vector<T>& values() const { return values_; }

 private:
  struct Slot {
    uint32_t version, dense_or_next;
  };
  std::vector<Slot> slots_;
  std::vector<Key> keys_;
  std::vector<T> values_;
  uint32_t free_head_ = 1;
};

// ---------------------------------------------------------------------------------------------
// NavigationMesh::validate (nav_mesh.rs:19-35, 187-407) and a single-island NavigationData.
// Island linking (boundary links, modified nodes, regions) is in landmass_b200_linking.hpp; animation links
// are mirrored in Python only.
// ---------------------------------------------------------------------------------------------
struct ValidationError : std::runtime_error {  // nav_mesh.rs:131-162
  enum Kind {
    TypeIndicesHaveWrongLength,
    NotEnoughVerticesInPolygon,
    InvalidVertexIndexInPolygon,
    DegenerateEdgeInPolygon,
    DoublyConnectedEdge,
    ConcavePolygon,
    IncorrectNumberOfPolygons,  // height mesh (nav_mesh.rs:410-479): expected, found
    InvalidPolygonIndices,      // height polygon index
    InvalidIndexInTriangle,     // triangle index, vertex index
    ClockwiseTriangle           // height polygon index
  } kind;
  size_t a, b;  // the variant's payload (polygon index / the two vertex indices / the two lengths)
  ValidationError(Kind k, size_t a_, size_t b_ = 0)
      : std::runtime_error("invalid navigation mesh (ValidationError kind " + std::to_string((int)k) + ")"),
        kind(k), a(a_), b(b_) {}
};

// nav_mesh.rs:483-583 as flat arrays (island-local, landmass coordinates), the layout lmb_navdata_desc takes
struct ValidNavigationMesh {
  std::vector<float> vertices;              // [V*3]
  std::vector<uint32_t> poly_edge_begin;    // [P+1]
  std::vector<uint32_t> poly_verts;         // [E] vertex of edge e's start (edge e runs vertex[e] -> vertex[e+1])
  std::vector<uint32_t> edge_neighbor;      // [E] polygon across the edge or LMB_NONE
  std::vector<uint32_t> edge_reverse;       // [E] index of the same edge inside the neighbour
  std::vector<uint32_t> poly_type, poly_region;  // [P]
  std::vector<float> poly_bounds;           // [P*6]
  std::vector<float> poly_center;           // [P*3]
  bool bounds_empty = true;
  std::array<float, 6> mesh_bounds{};       // min xyz, max xyz
  std::vector<std::pair<uint32_t, uint32_t>> boundary_edges;  // (polygon, edge), ascending
  // optional height mesh (nav_mesh.rs:503-521): per polygon {base vertex, vertex count, base triangle, triangle count}
  bool has_height = false;
  std::vector<uint32_t> hpoly;   // [P*4]
  std::vector<float> hverts;     // [HV*3]
  std::vector<uint8_t> htris;    // [HT*3] indices relative to the polygon's base vertex
  size_t n_polys() const { return poly_type.size(); }
};

struct HeightPolygon {  // nav_mesh.rs:74-90
  uint32_t base_vertex_index, vertex_count, base_triangle_index, triangle_count;
};
struct HeightNavigationMesh {  // nav_mesh.rs:52-68
  std::vector<HeightPolygon> polygons;
  std::vector<Vec3> vertices;
  std::vector<std::array<uint32_t, 3>> triangles;
};

struct NavigationMesh {  // nav_mesh.rs:19-35 (vertices already in landmass coordinates; polygons counter-clockwise)
  std::vector<Vec3> vertices;
  std::vector<std::vector<uint32_t>> polygons;
  std::vector<uint32_t> polygon_type_indices;
  std::optional<HeightNavigationMesh> height_mesh;

  ValidNavigationMesh validate() const {
    using VE = ValidationError;
    const size_t P = polygons.size(), V = vertices.size();
    if (P != polygon_type_indices.size()) throw VE(VE::TypeIndicesHaveWrongLength, P, polygon_type_indices.size());
    ValidNavigationMesh m;
    if (height_mesh) {  // HeightNavigationMesh::validate (nav_mesh.rs:410-479), before the polygons are looked at
      const HeightNavigationMesh& h = *height_mesh;
      if (h.polygons.size() != P) throw VE(VE::IncorrectNumberOfPolygons, P, h.polygons.size());
      for (size_t pi = 0; pi < P; pi++) {
        const HeightPolygon& hp = h.polygons[pi];
        const size_t last = (size_t)hp.base_triangle_index + hp.triangle_count;
        if (last > h.triangles.size()) throw VE(VE::InvalidPolygonIndices, pi);
        for (size_t ti = hp.base_triangle_index; ti < last; ti++) {
          size_t real[3];
          for (int k = 0; k < 3; k++) {
            const size_t r = (size_t)hp.base_vertex_index + h.triangles[ti][k];
            if (r >= h.vertices.size() || h.triangles[ti][k] >= hp.vertex_count) throw VE(VE::InvalidIndexInTriangle, ti, r);
            real[k] = r;
          }
          const Vec3 &a = h.vertices[real[0]], &b = h.vertices[real[1]], &c = h.vertices[real[2]];
          const float abx = b[0] - a[0], aby = b[1] - a[1], acx = c[0] - a[0], acy = c[1] - a[1];
          const float lhs = abx * acy, rhs = aby * acx;
          if (lhs - rhs < 0.0f) throw VE(VE::ClockwiseTriangle, pi);
        }
        const uint32_t q[4] = {hp.base_vertex_index, hp.vertex_count, hp.base_triangle_index, hp.triangle_count};
        m.hpoly.insert(m.hpoly.end(), q, q + 4);
      }
      m.has_height = true;
      for (const Vec3& v : h.vertices) m.hverts.insert(m.hverts.end(), v.begin(), v.end());
      for (const auto& t : h.triangles)
        for (int k = 0; k < 3; k++) m.htris.push_back((uint8_t)t[k]);
    }
    for (const Vec3& v : vertices) m.vertices.insert(m.vertices.end(), v.begin(), v.end());
    m.poly_edge_begin.assign(P + 1, 0u);
    // the reference discovers errors polygon by polygon: size, vertex range, then per edge degenerate /
    // doubly connected / concave (nav_mesh.rs:246-319)
    struct EdgeUse {
      uint32_t polygon, edge;
    };
    struct EdgeSlot {
      uint64_t key;
      EdgeUse first;
      bool has_second = false;
      EdgeUse second{};
    };
    std::vector<EdgeSlot> table;  // open addressing on the (lo, hi) vertex pair
    size_t total_edges = 0;
    for (const auto& poly : polygons) total_edges += poly.size();
    size_t cap = 16;
    while (cap < 2 * total_edges + 2) cap <<= 1;
    table.assign(cap, EdgeSlot{~0ull, {0, 0}, false, {0, 0}});
    auto slot_of = [&](uint64_t key) -> EdgeSlot& {
      size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 17) & (cap - 1);
      while (table[h].key != ~0ull && table[h].key != key) h = (h + 1) & (cap - 1);
      return table[h];
    };
    for (size_t p = 0; p < P; p++) {
      const auto& poly = polygons[p];
      const size_t n = poly.size();
      if (n < 3) throw VE(VE::NotEnoughVerticesInPolygon, p);
      for (uint32_t v : poly)
        if (v >= V) throw VE(VE::InvalidVertexIndexInPolygon, p);
      for (size_t e = 0; e < n; e++) {
        const uint32_t a = poly[e], b = poly[(e + 1) % n], l = poly[(e + n - 1) % n];
        if (a == b) throw VE(VE::DegenerateEdgeInPolygon, p);
        const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
        EdgeSlot& s = slot_of(((uint64_t)lo << 32) | hi);
        if (s.key == ~0ull) {
          s.key = ((uint64_t)lo << 32) | hi;
          s.first = {(uint32_t)p, (uint32_t)e};
        } else if (!s.has_second) {
          s.has_second = true;
          s.second = {(uint32_t)p, (uint32_t)e};
        } else {
          throw VE(VE::DoublyConnectedEdge, lo, hi);
        }
        // convexity at vertex a (nav_mesh.rs:300-319), f32
        const float lx = vertices[l][0] - vertices[a][0], ly = vertices[l][1] - vertices[a][1];
        const float rx = vertices[b][0] - vertices[a][0], ry = vertices[b][1] - vertices[a][1];
        const float pd = rx * ly - ry * lx, dt = rx * lx + ry * ly;
        if (!(pd > 0.0f || (pd == 0.0f && dt < 0.0f))) throw VE(VE::ConcavePolygon, p);
      }
      m.poly_edge_begin[p + 1] = m.poly_edge_begin[p] + (uint32_t)n;
      m.poly_verts.insert(m.poly_verts.end(), poly.begin(), poly.end());
    }
    const size_t E = m.poly_verts.size();
    m.edge_neighbor.assign(E, LMB_NONE);
    m.edge_reverse.assign(E, 0u);
    for (const EdgeSlot& s : table) {
      if (s.key == ~0ull || !s.has_second) continue;
      const uint32_t e1 = m.poly_edge_begin[s.first.polygon] + s.first.edge;
      const uint32_t e2 = m.poly_edge_begin[s.second.polygon] + s.second.edge;
      m.edge_neighbor[e1] = s.second.polygon, m.edge_reverse[e1] = s.second.edge;
      m.edge_neighbor[e2] = s.first.polygon, m.edge_reverse[e2] = s.first.edge;
    }
    // regions: connected components numbered by first appearance (nav_mesh.rs:354-364)
    m.poly_region.assign(P, LMB_NONE);
    m.poly_type = polygon_type_indices;
    uint32_t next_region = 0;
    std::vector<uint32_t> stack;
    for (size_t p = 0; p < P; p++) {
      if (m.poly_region[p] != LMB_NONE) continue;
      m.poly_region[p] = next_region;
      stack.push_back((uint32_t)p);
      while (!stack.empty()) {
        const uint32_t q = stack.back();
        stack.pop_back();
        for (uint32_t e = m.poly_edge_begin[q]; e < m.poly_edge_begin[q + 1]; e++) {
          const uint32_t nb = m.edge_neighbor[e];
          if (nb != LMB_NONE && m.poly_region[nb] == LMB_NONE) {
            m.poly_region[nb] = next_region;
            stack.push_back(nb);
          }
        }
      }
      next_region++;
    }
    // bounds and centres (nav_mesh.rs:321-352): centre = (sum of the vertices in order) / n, f32
    m.poly_bounds.resize(P * 6), m.poly_center.resize(P * 3);
    for (size_t p = 0; p < P; p++) {
      float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, sum[3] = {0, 0, 0};
      for (uint32_t v : polygons[p])
        for (int k = 0; k < 3; k++) {
          lo[k] = std::fmin(lo[k], vertices[v][k]);
          hi[k] = std::fmax(hi[k], vertices[v][k]);
          sum[k] = sum[k] + vertices[v][k];
        }
      for (int k = 0; k < 3; k++) {
        m.poly_bounds[6 * p + k] = lo[k];
        m.poly_bounds[6 * p + 3 + k] = hi[k];
        m.poly_center[3 * p + k] = sum[k] / (float)polygons[p].size();
      }
    }
    // with a height mesh the bounds come from its vertices (nav_mesh.rs:334-345)
    const std::vector<Vec3>& bsrc = height_mesh ? height_mesh->vertices : vertices;
    if (height_mesh)
      for (size_t p = 0; p < P; p++) {
        const HeightPolygon& hp = height_mesh->polygons[p];
        float lo[3] = {1.0f, 1.0f, 1.0f}, hi[3] = {0.0f, 0.0f, 0.0f};  // Empty
        if (hp.vertex_count) {
          for (int k = 0; k < 3; k++) lo[k] = INFINITY, hi[k] = -INFINITY;
          for (uint32_t v = hp.base_vertex_index; v < hp.base_vertex_index + hp.vertex_count; v++)
            for (int k = 0; k < 3; k++) lo[k] = std::fmin(lo[k], bsrc[v][k]), hi[k] = std::fmax(hi[k], bsrc[v][k]);
        }
        for (int k = 0; k < 3; k++) m.poly_bounds[6 * p + k] = lo[k], m.poly_bounds[6 * p + 3 + k] = hi[k];
      }
    m.bounds_empty = bsrc.empty();
    if (!bsrc.empty()) {
      for (int k = 0; k < 3; k++) m.mesh_bounds[k] = INFINITY, m.mesh_bounds[3 + k] = -INFINITY;
      for (const Vec3& v : bsrc)
        for (int k = 0; k < 3; k++) {
          m.mesh_bounds[k] = std::fmin(m.mesh_bounds[k], v[k]);
          m.mesh_bounds[3 + k] = std::fmax(m.mesh_bounds[3 + k], v[k]);
        }
    }
    for (size_t p = 0; p < P; p++)
      for (uint32_t e = m.poly_edge_begin[p]; e < m.poly_edge_begin[p + 1]; e++)
        if (m.edge_neighbor[e] == LMB_NONE) m.boundary_edges.emplace_back((uint32_t)p, e - m.poly_edge_begin[p]);
    return m;
  }
};

// util.rs:315-368 / island.rs:16-27: one island = a transform and a validated mesh
struct Transform {
  Vec3 translation{};
  float rotation = 0.0f;
};

// The flattened navigation data of ONE island (no links, no modified nodes, no region numbers: exactly what
// NavigationData::flatten yields for such a world).  Owns the arrays `desc` points into.
class SingleIslandNavigationData {
 public:
  SingleIslandNavigationData(Transform transform, ValidNavigationMesh mesh,
                             std::vector<std::pair<uint32_t, float>> type_index_costs = {})
      : mesh_(std::move(mesh)) {
    translation_ = transform.translation;
    rotation_ = transform.rotation;
    const uint32_t P = (uint32_t)mesh_.n_polys(), V = (uint32_t)(mesh_.vertices.size() / 3);
    if (mesh_.bounds_empty)
      bounds_ = {1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f};
    else
      bounds_ = mesh_.mesh_bounds;
    poly_begin_[0] = 0, poly_begin_[1] = P, vert_begin_[0] = 0, vert_begin_[1] = V;
    for (size_t x = 1; x < type_index_costs.size(); x++)  // ascending type index
      for (size_t y = x; y > 0 && type_index_costs[y].first < type_index_costs[y - 1].first; y--)
        std::swap(type_index_costs[y], type_index_costs[y - 1]);
    for (const auto& kv : type_index_costs) cost_index_.push_back(kv.first), cost_value_.push_back(kv.second);
    node_link_begin_.assign((size_t)P + 1, 0u);
    node_region_number_.assign(P, LMB_NONE);
    lmb_navdata_desc& d = desc_;
    d = lmb_navdata_desc{};
    d.struct_size = (uint32_t)sizeof(lmb_navdata_desc);
    d.n_islands = 1;
    d.island_translation = translation_.data(), d.island_rotation = &rotation_, d.island_bounds = bounds_.data();
    d.island_poly_begin = poly_begin_, d.island_vert_begin = vert_begin_;
    d.n_vertices = V, d.vertices = mesh_.vertices.data();
    d.n_polys = P, d.poly_edge_begin = mesh_.poly_edge_begin.data();
    d.n_edges = (uint32_t)mesh_.poly_verts.size();
    d.edge_vertex = mesh_.poly_verts.data(), d.edge_neighbor = mesh_.edge_neighbor.data();
    d.edge_reverse = mesh_.edge_reverse.data();
    d.poly_type = mesh_.poly_type.data(), d.poly_region = mesh_.poly_region.data();
    d.poly_bounds = mesh_.poly_bounds.data(), d.poly_center = mesh_.poly_center.data();
    d.n_type_costs = (uint32_t)cost_index_.size();
    d.type_cost_index = cost_index_.data(), d.type_cost_value = cost_value_.data();
    d.node_link_begin = node_link_begin_.data();
    if (mesh_.has_height) {
      has_height_ = 1;
      d.island_has_height = &has_height_, d.hpoly = mesh_.hpoly.data();
      d.n_hverts = (uint32_t)(mesh_.hverts.size() / 3), d.hverts = mesh_.hverts.data();
      d.n_htris = (uint32_t)(mesh_.htris.size() / 3), d.htris = mesh_.htris.data();
    }
    d.modified_edge_begin = zero2_, d.modified_vert_begin = zero2_;
    d.node_region_number = node_region_number_.data();
  }
  SingleIslandNavigationData(const SingleIslandNavigationData&) = delete;
  SingleIslandNavigationData& operator=(const SingleIslandNavigationData&) = delete;
  const lmb_navdata_desc& desc() const { return desc_; }
  const ValidNavigationMesh& mesh() const { return mesh_; }

 private:
  ValidNavigationMesh mesh_;
  Vec3 translation_{};
  float rotation_ = 0.0f;
  std::array<float, 6> bounds_{};
  uint32_t poly_begin_[2]{}, vert_begin_[2]{}, zero2_[2] = {0u, 0u};
  uint8_t has_height_ = 0;
  std::vector<uint32_t> cost_index_, node_link_begin_, node_region_number_;
  std::vector<float> cost_value_;
  lmb_navdata_desc desc_{};
};

// query.rs:19-55: a point on the navigation meshes (`node` is the global node id of the C ABI)
struct SampledPoint {
  Vec3 point;
  uint32_t node;
};
// query.rs:105-127
struct PathStep {
  bool is_animation_link = false;
  Vec3 point{};      // Waypoint, or the animation link's start point
  Vec3 end_point{};  // the animation link's end point
  uint32_t link_id = LMB_NONE;
};
struct FindPathError : std::runtime_error {  // query.rs:97-103 (NoPathFound is reported as an empty optional)
  uint32_t type_index;
  float cost;
  FindPathError(uint32_t t, float c)
      : std::runtime_error("NonPositiveTypeIndexCost"), type_index(t), cost(c) {}
};

// lib.rs:538-547
struct PathingResult {
  AgentId agent;
  bool success;
  uint32_t explored_nodes;
};

struct Error : std::runtime_error {
  int32_t status;
  Error(int32_t st, const std::string& what) : std::runtime_error(what), status(st) {}
};

// The entry points `update` needs.  `Ctx` is the backend's opaque context type.
struct CudaBackend {
  using Ctx = lmb_ctx;
  static int32_t navdata_upload(Ctx* c, const lmb_navdata_desc* d) { return lmb_navdata_upload(c, d); }
  static int32_t navdata_update(Ctx* c, const lmb_navdata_desc* d, const lmb_nav_remap* r) {
    return lmb_navdata_update(c, d, r);
  }
  static int32_t step(Ctx* c, const lmb_agents_in* a, const lmb_characters_in* ch, const lmb_options* o, float dt,
                      const lmb_agents_out* out) {
    return lmb_step(c, a, ch, o, dt, out);
  }
  static int32_t pathing_results(Ctx* c, lmb_pathing_result* out, uint32_t cap, uint32_t* n) {
    return lmb_pathing_results(c, out, cap, n);
  }
  static int32_t sample_points(Ctx* c, uint32_t n, const float* points, const lmb_options* o, float* out_points,
                               uint32_t* out_nodes) {
    return lmb_sample_points(c, n, points, o, out_points, out_nodes);
  }
  static int32_t agent_path(Ctx* c, uint32_t slot, uint32_t* nodes, uint32_t* portals, uint32_t cap, uint32_t* len) {
    return lmb_agent_path(c, slot, nodes, portals, cap, len);
  }
  static int32_t agent_avoidance_constraints(Ctx* c, uint32_t agent_index, uint32_t cap, float* lines, uint32_t* n_orig,
                                             uint32_t* n_fallback, uint32_t* ran) {
    return lmb_agent_avoidance_constraints(c, agent_index, cap, lines, n_orig, n_fallback, ran);
  }
  static int32_t find_straight_paths(Ctx* c, uint32_t n, const uint32_t* sn, const float* sp, const uint32_t* en,
                                     const float* ep, const uint32_t* ob, const uint32_t* ot, const float* oc,
                                     uint32_t* succ, uint32_t* begin, uint32_t* kind, float* pts, uint32_t* link,
                                     uint32_t cap, const uint32_t* pa, const uint32_t* pb, const uint32_t* pk) {
    return lmb_find_straight_paths(c, n, sn, sp, en, ep, ob, ot, oc, succ, begin, kind, pts, link, cap, pa, pb, pk);
  }
  static std::string last_error(Ctx* c) {
    const char* e = lmb_last_error(c);
    return e ? e : "";
  }
};

// `Archipelago` (lib.rs:49-61) over a context the caller created (lmb_create) and destroys.
template <class Backend>
class BasicArchipelago {
 public:
  ArchipelagoOptions archipelago_options;

  BasicArchipelago(typename Backend::Ctx* ctx, ArchipelagoOptions options, uint32_t max_agents)
      : archipelago_options(options), ctx_(ctx), max_agents_(max_agents) {}

  // the flattened navigation data (NavigationData::flatten); every island counts as changed
  void set_navigation_data(const lmb_navdata_desc& desc) { check(Backend::navdata_upload(ctx_, &desc), "lmb_navdata_upload"); }

  AgentId add_agent(Agent agent) {  // lib.rs:96-101
    const AgentId id = agents_.insert(std::move(agent));
    if (id.idx >= max_agents_) {
      agents_.remove(id);
      throw Error(LMB_ERR_CAPACITY, "more live agents than the context's max_agents");
    }
    return id;
  }
  void remove_agent(AgentId id) {  // lib.rs:103-108: panics on a stale id
    if (!agents_.remove(id)) throw std::out_of_range("Agent should be present in the archipelago");
  }
  Agent* get_agent_mut(AgentId id) { return agents_.get(id); }
  const Agent* get_agent(AgentId id) const { return agents_.get(id); }
  const std::vector<AgentId>& get_agent_ids() const { return agents_.keys(); }

  CharacterId add_character(Character c) { return characters_.insert(c); }  // lib.rs:125-128
  void remove_character(CharacterId id) {
    if (!characters_.remove(id)) throw std::out_of_range("Character should be present in the archipelago");
  }
  Character* get_character_mut(CharacterId id) { return characters_.get(id); }
  const std::vector<CharacterId>& get_character_ids() const { return characters_.keys(); }

  const std::vector<PathingResult>& get_pathing_results() const { return pathing_results_; }  // lib.rs:233-236

  // `Archipelago::sample_point` (lib.rs:235-246, query.rs:70-94); an empty optional is SamplePointError::OutOfRange.
  // `distances` replaces the archipelago's point_sample_distance for this call (only its first four fields are read).
  std::optional<SampledPoint> sample_point(Vec3 point, const lmb_options* distances = nullptr) {
    lmb_options o = archipelago_options.abi;
    if (distances) {
      o.horizontal_distance = distances->horizontal_distance, o.distance_above = distances->distance_above;
      o.distance_below = distances->distance_below, o.vertical_preference_ratio = distances->vertical_preference_ratio;
    }
    float out_point[3];
    uint32_t node = LMB_NONE;
    check(Backend::sample_points(ctx_, 1, point.data(), &o, out_point, &node), "lmb_sample_points");
    if (node == LMB_NONE) return std::nullopt;
    return SampledPoint{{out_point[0], out_point[1], out_point[2]}, node};
  }

  // `Archipelago::find_path` (lib.rs:248-268, query.rs:185-273): the straight path between two sampled points;
  // an empty optional is FindPathError::NoPathFound.
  std::optional<std::vector<PathStep>> find_path(const SampledPoint& start, const SampledPoint& end,
                                                 std::vector<std::pair<uint32_t, float>> override_type_index_costs = {},
                                                 const PermittedAnimationLinks& permitted = {}) {
    for (const auto& kv : override_type_index_costs)
      if (!(kv.second > 0.0f)) throw FindPathError(kv.first, kv.second);
    for (size_t x = 1; x < override_type_index_costs.size(); x++)
      for (size_t y = x; y > 0 && override_type_index_costs[y].first < override_type_index_costs[y - 1].first; y--)
        std::swap(override_type_index_costs[y], override_type_index_costs[y - 1]);
    std::vector<uint32_t> ot, kinds = permitted.kinds;
    std::vector<float> oc;
    for (const auto& kv : override_type_index_costs) ot.push_back(kv.first), oc.push_back(kv.second);
    sort_unique(kinds);
    const uint32_t ob[2] = {0u, (uint32_t)ot.size()}, pb[2] = {0u, (uint32_t)kinds.size()};
    const uint32_t permit_all = permitted.all ? 1u : 0u;
    uint32_t cap = 256;
    for (;;) {
      std::vector<uint32_t> kind(cap), link(cap);
      std::vector<float> pts(6 * (size_t)cap);
      uint32_t success = 0, begin[2] = {0, 0};
      const int32_t st = Backend::find_straight_paths(ctx_, 1, &start.node, start.point.data(), &end.node,
                                                      end.point.data(), ob, ot.data(), oc.data(), &success, begin,
                                                      kind.data(), pts.data(), link.data(), cap, &permit_all, pb,
                                                      kinds.data());
      if (st == LMB_ERR_CAPACITY) {
        cap = begin[1] > cap ? begin[1] : cap * 2;
        continue;
      }
      check(st, "lmb_find_straight_paths");
      if (!success) return std::nullopt;
      std::vector<PathStep> steps;
      for (uint32_t k = begin[0]; k < begin[1]; k++) {
        PathStep ps;
        ps.is_animation_link = kind[k] == 1;
        for (int c = 0; c < 3; c++) ps.point[c] = pts[6 * (size_t)k + c], ps.end_point[c] = pts[6 * (size_t)k + 3 + c];
        ps.link_id = ps.is_animation_link ? link[k] : LMB_NONE;
        steps.push_back(ps);
      }
      // the sampled points are returned as given (query.rs:224, 227)
      if (!steps.empty()) {
        steps.front().point = start.point;
        if (!steps.back().is_animation_link) steps.back().point = end.point;
      }
      return steps;
    }
  }

  // `agent.current_path` read-back for debug drawing (debug.rs:300-350): the corridor as flat (node id, portal code)
  // pairs — portal = edge index, 0x80000000 | link index, or LMB_NONE behind the last node; empty = no path.
  std::pair<std::vector<uint32_t>, std::vector<uint32_t>> agent_path(AgentId id) {
    std::vector<uint32_t> nodes(64), portals(64);
    uint32_t len = 0;
    for (;;) {
      const int32_t st = Backend::agent_path(ctx_, id.idx, nodes.data(), portals.data(), (uint32_t)nodes.size(), &len);
      if (st == LMB_ERR_CAPACITY) {
        nodes.resize(len), portals.resize(len);
        continue;
      }
      check(st, "lmb_agent_path");
      break;
    }
    nodes.resize(len), portals.resize(len);
    return {nodes, portals};
  }

  // `Archipelago::update` (lib.rs:270-534): pack -> lmb_step -> scatter
  void update(float delta_time) {
    const std::vector<AgentId>& keys = agents_.keys();
    std::vector<Agent>& ag = agents_.values();
    const uint32_t n = (uint32_t)ag.size();
    slot_.resize(n), flags_.resize(n), radius_.resize(n), desired_speed_.resize(n), max_speed_.resize(n);
    reach_condition_.resize(n), reach_distance_.resize(n), anim_reach_.resize(n);
    position_.resize(3 * (size_t)n), velocity_.resize(3 * (size_t)n), target_.assign(3 * (size_t)n, 0.0f);
    override_begin_.assign((size_t)n + 1, 0u), permitted_begin_.assign((size_t)n + 1, 0u);
    override_type_.clear(), override_cost_.clear(), permitted_kind_.clear();
    if (device_version_.size() < max_agents_) device_version_.assign(max_agents_, 0u);
    std::vector<uint32_t> live(n);
    for (uint32_t i = 0; i < n; i++) {
      const Agent& a = ag[i];
      uint32_t f = 0;
      // a slot that was vacated and re-used carries another agent's corridor on the device
      if (device_version_[keys[i].idx] != keys[i].version) f |= LMB_AGENT_RESET;
      live[i] = keys[i].idx;
      if (a.current_target) {
        f |= LMB_AGENT_HAS_TARGET;
        for (int k = 0; k < 3; k++) target_[3 * (size_t)i + k] = (*a.current_target)[k];
      }
      if (a.paused) f |= LMB_AGENT_PAUSED;
      if (a.using_animation_link_) f |= LMB_AGENT_USING_ANIMATION_LINK;
      // (an agent that stopped asking still has to learn whether avoidance ran, to drop its old data)
      if (a.keep_avoidance_data || a.has_avoidance_data_) f |= LMB_AGENT_KEEP_AVOIDANCE_DATA;
      if (a.permitted_animation_links.all) {
        f |= LMB_AGENT_PERMIT_ALL_LINKS;
      } else {
        std::vector<uint32_t> kinds = a.permitted_animation_links.kinds;
        sort_unique(kinds);
        permitted_kind_.insert(permitted_kind_.end(), kinds.begin(), kinds.end());
      }
      permitted_begin_[i + 1] = (uint32_t)permitted_kind_.size();
      slot_[i] = keys[i].idx;
      flags_[i] = f;
      for (int k = 0; k < 3; k++) {
        position_[3 * (size_t)i + k] = a.position[k];
        velocity_[3 * (size_t)i + k] = a.velocity[k];
      }
      radius_[i] = a.radius;
      desired_speed_[i] = a.desired_speed;
      max_speed_[i] = a.max_speed;
      reach_condition_[i] = a.target_reached_condition.kind;
      reach_distance_[i] = a.target_reached_condition.distance ? *a.target_reached_condition.distance : -1.0f;
      anim_reach_[i] = a.animation_link_reached_distance ? *a.animation_link_reached_distance : -1.0f;
      // HashMap in the reference: ascending type index is the canonical order of the ABI
      std::vector<std::pair<uint32_t, float>> ov = a.overrides_;
      for (size_t x = 1; x < ov.size(); x++)
        for (size_t y = x; y > 0 && ov[y].first < ov[y - 1].first; y--) std::swap(ov[y], ov[y - 1]);
      for (const auto& kv : ov) {
        override_type_.push_back(kv.first);
        override_cost_.push_back(kv.second);
      }
      override_begin_[i + 1] = (uint32_t)override_type_.size();
    }
    // versions of the slots that are alive now; everything else is forgotten (a later re-use resets)
    std::vector<uint32_t> versions(max_agents_, 0u);
    for (uint32_t i = 0; i < n; i++) versions[live[i]] = keys[i].version;
    device_version_.swap(versions);

    const std::vector<Character>& ch = characters_.values();
    const uint32_t nc = (uint32_t)ch.size();
    char_position_.resize(3 * (size_t)nc), char_velocity_.resize(3 * (size_t)nc), char_radius_.resize(nc);
    for (uint32_t i = 0; i < nc; i++) {
      for (int k = 0; k < 3; k++) {
        char_position_[3 * (size_t)i + k] = ch[i].position[k];
        char_velocity_[3 * (size_t)i + k] = ch[i].velocity[k];
      }
      char_radius_[i] = ch[i].radius;
    }

    lmb_agents_in in{};
    in.n = n;
    in.slot = slot_.data(), in.flags = flags_.data();
    in.position = position_.data(), in.velocity = velocity_.data(), in.target = target_.data();
    in.radius = radius_.data(), in.desired_speed = desired_speed_.data(), in.max_speed = max_speed_.data();
    in.reach_condition = reach_condition_.data(), in.reach_distance = reach_distance_.data();
    in.anim_link_reach_distance = anim_reach_.data();
    in.override_begin = override_begin_.data(), in.override_type = override_type_.data();
    in.override_cost = override_cost_.data();
    in.permitted_begin = permitted_begin_.data(), in.permitted_kind = permitted_kind_.data();
    lmb_characters_in cin{};
    cin.n = nc;
    cin.position = char_position_.data(), cin.velocity = char_velocity_.data(), cin.radius = char_radius_.data();
    out_velocity_.resize(3 * (size_t)n + 1), out_state_.resize((size_t)n + 1), out_link_id_.resize((size_t)n + 1);
    out_link_start_.resize(3 * (size_t)n + 1), out_link_end_.resize(3 * (size_t)n + 1);
    lmb_agents_out out{};
    out.desired_velocity = out_velocity_.data(), out.state = out_state_.data();
    out.reached_link_id = out_link_id_.data(), out.reached_link_start = out_link_start_.data();
    out.reached_link_end = out_link_end_.data();
    check(Backend::step(ctx_, &in, &cin, &archipelago_options.abi, delta_time, &out), "lmb_step");

    for (uint32_t i = 0; i < n; i++) {
      Agent& a = ag[i];
      a.current_desired_move_ = {out_velocity_[3 * (size_t)i], out_velocity_[3 * (size_t)i + 1],
                                 out_velocity_[3 * (size_t)i + 2]};
      a.state_ = (AgentState)out_state_[i];
      if (out_link_id_[i] != LMB_NONE) {
        ReachedAnimationLink r;
        r.link_id = out_link_id_[i];
        for (int k = 0; k < 3; k++) {
          r.start_point[k] = out_link_start_[3 * (size_t)i + k];
          r.end_point[k] = out_link_end_[3 * (size_t)i + k];
        }
        a.current_animation_link_ = r;
      } else {
        a.current_animation_link_.reset();
      }
      if (a.keep_avoidance_data || a.has_avoidance_data_) {
        // agent.avoidance_data = agent.keep_avoidance_data.then_some(debug_data), only for the agents avoidance
        // visited (avoidance.rs:83-87, 161-172)
        uint32_t n_orig = 0, n_fb = 0, ran = 0, cap = 64;
        std::vector<float> lines;
        for (;;) {
          lines.resize(4 * (size_t)cap);
          const int32_t st = Backend::agent_avoidance_constraints(ctx_, i, cap, lines.data(), &n_orig, &n_fb, &ran);
          if (st == LMB_ERR_CAPACITY) {
            cap = n_orig + n_fb;
            continue;
          }
          check(st, "lmb_agent_avoidance_constraints");
          break;
        }
        if (ran) {
          if (!a.keep_avoidance_data) {
            a.has_avoidance_data_ = false;
            a.avoidance_data_ = Agent::AvoidanceData{};
          } else {
            Agent::AvoidanceData d;
            d.fell_back = ran == 2;
            for (uint32_t k = 0; k < n_orig + n_fb; k++) {
              const std::array<float, 4> l{lines[4 * (size_t)k], lines[4 * (size_t)k + 1], lines[4 * (size_t)k + 2],
                                           lines[4 * (size_t)k + 3]};
              (k < n_orig ? d.original : d.fallback).push_back(l);
            }
            a.avoidance_data_ = std::move(d);
            a.has_avoidance_data_ = true;
          }
        }
      }
    }
    // lib.rs:406-410: one result per agent that re-planned, in agent order
    std::vector<lmb_pathing_result> raw((size_t)n + 1);
    uint32_t n_results = 0;
    check(Backend::pathing_results(ctx_, raw.data(), (uint32_t)raw.size(), &n_results), "lmb_pathing_results");
    pathing_results_.clear();
    for (uint32_t k = 0; k < n_results; k++)
      pathing_results_.push_back({keys[raw[k].agent], raw[k].success != 0, raw[k].explored_nodes});
  }

 private:
  static void sort_unique(std::vector<uint32_t>& v) {
    for (size_t x = 1; x < v.size(); x++)
      for (size_t y = x; y > 0 && v[y] < v[y - 1]; y--) std::swap(v[y], v[y - 1]);
    size_t w = 0;
    for (size_t x = 0; x < v.size(); x++)
      if (w == 0 || v[x] != v[w - 1]) v[w++] = v[x];
    v.resize(w);
  }
  void check(int32_t st, const char* what) {
    if (st != LMB_OK) throw Error(st, std::string(what) + ": " + Backend::last_error(ctx_));
  }

  typename Backend::Ctx* ctx_;
  uint32_t max_agents_;
  DenseSlotMap<Agent> agents_;
  DenseSlotMap<Character> characters_;
  std::vector<PathingResult> pathing_results_;
  std::vector<uint32_t> device_version_;
  // packed per-frame arrays (kept between frames: no allocation in steady state)
  std::vector<uint32_t> slot_, flags_, reach_condition_, override_begin_, override_type_, permitted_begin_,
      permitted_kind_, out_state_, out_link_id_;
  std::vector<float> position_, velocity_, target_, radius_, desired_speed_, max_speed_, reach_distance_, anim_reach_,
      override_cost_, char_position_, char_velocity_, char_radius_, out_velocity_, out_link_start_, out_link_end_;
};

using Archipelago = BasicArchipelago<CudaBackend>;

}  // namespace landmass_b200
