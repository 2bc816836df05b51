// TEST INFRASTRUCTURE — CPU oracle for landmass_b200 (see oracle/README.md).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call anything under oracle/.
//
// Single-threaded C++17 restatement of the reference's per-frame agent step,
// `Archipelago::update` (crates/landmass/src/lib.rs:270-534), phases A-C here;
// phase D (avoidance) lives in avoidance_oracle.cpp.  The real crate cannot be
// built in this environment (no Rust toolchain), so this restatement is pinned
// by the reference's own known-answer tests transcribed under tests/golden/.
//
// The data structures deliberately mirror the reference's algorithmic
// complexity (linear polygon scan for sampling, BinaryHeap + HashMap A*), so
// that timing this code is a fair stand-in for the reference's CPU path.
//
// Compile with: g++ -O2 -std=c++17 -ffp-contract=off (no FMA contraction).
#include "landmass_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <optional>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "orc_internal.h"

namespace orc {

// ---------------------------------------------------------------------------
// NavData: copy of the flattened description.
// ---------------------------------------------------------------------------
template <class T>
static std::vector<T> copy_arr(const T* p, size_t n) {
  if (!p || n == 0) return {};
  return std::vector<T>(p, p + n);
}

void NavData::load(const lmb_navdata_desc& d) {
  n_islands = d.n_islands;
  islands.clear();
  for (uint32_t i = 0; i < d.n_islands; i++) {
    Island is;
    is.xf.translation = {d.island_translation[3 * i], d.island_translation[3 * i + 1],
                         d.island_translation[3 * i + 2]};
    is.xf.rotation = d.island_rotation[i];
    is.bmin = {d.island_bounds[6 * i], d.island_bounds[6 * i + 1], d.island_bounds[6 * i + 2]};
    is.bmax = {d.island_bounds[6 * i + 3], d.island_bounds[6 * i + 4], d.island_bounds[6 * i + 5]};
    is.bounds_empty = !(is.bmin.x <= is.bmax.x);
    is.poly_begin = d.island_poly_begin[i];
    is.poly_end = d.island_poly_begin[i + 1];
    is.vert_begin = d.island_vert_begin[i];
    is.vert_end = d.island_vert_begin[i + 1];
    is.has_height = d.island_has_height ? d.island_has_height[i] != 0 : false;
    islands.push_back(is);
  }
  vertices.resize(d.n_vertices);
  for (uint32_t i = 0; i < d.n_vertices; i++)
    vertices[i] = {d.vertices[3 * i], d.vertices[3 * i + 1], d.vertices[3 * i + 2]};
  n_polys = d.n_polys;
  poly_edge_begin = copy_arr(d.poly_edge_begin, d.n_polys + 1);
  edge_vertex = copy_arr(d.edge_vertex, d.n_edges);
  edge_neighbor = copy_arr(d.edge_neighbor, d.n_edges);
  edge_reverse = copy_arr(d.edge_reverse, d.n_edges);
  poly_type = copy_arr(d.poly_type, d.n_polys);
  poly_region = copy_arr(d.poly_region, d.n_polys);
  poly_bounds = copy_arr(d.poly_bounds, (size_t)d.n_polys * 6);
  poly_center = copy_arr(d.poly_center, (size_t)d.n_polys * 3);
  poly_island.assign(d.n_polys, 0);
  for (uint32_t i = 0; i < d.n_islands; i++)
    for (uint32_t p = islands[i].poly_begin; p < islands[i].poly_end; p++) poly_island[p] = i;
  hpoly = copy_arr(d.hpoly, d.hpoly ? (size_t)d.n_polys * 4 : 0);
  hverts.resize(d.hverts ? d.n_hverts : 0);
  for (size_t i = 0; i < hverts.size(); i++)
    hverts[i] = {d.hverts[3 * i], d.hverts[3 * i + 1], d.hverts[3 * i + 2]};
  htris = copy_arr(d.htris, d.htris ? (size_t)d.n_htris * 3 : 0);
  type_cost.clear();
  for (uint32_t i = 0; i < d.n_type_costs; i++) type_cost[d.type_cost_index[i]] = d.type_cost_value[i];
  n_links = d.n_links;
  link_src_node = copy_arr(d.link_src_node, d.n_links);
  link_dst_node = copy_arr(d.link_dst_node, d.n_links);
  link_dst_type = copy_arr(d.link_dst_type, d.n_links);
  link_kind = copy_arr(d.link_kind, d.n_links);
  link_reverse = copy_arr(d.link_reverse, d.n_links);
  link_cost = copy_arr(d.link_cost, d.n_links);
  link_anim_kind = copy_arr(d.link_anim_kind, d.n_links);
  link_anim_id = copy_arr(d.link_anim_id, d.n_links);
  link_portal.resize(d.n_links);
  link_dst_portal.resize(d.n_links);
  for (uint32_t i = 0; i < d.n_links; i++) {
    const float* p = d.link_portal + 6 * i;
    link_portal[i] = {{p[0], p[1], p[2]}, {p[3], p[4], p[5]}};
    if (d.link_dst_portal) {
      const float* q = d.link_dst_portal + 6 * i;
      link_dst_portal[i] = {{q[0], q[1], q[2]}, {q[3], q[4], q[5]}};
    } else {
      link_dst_portal[i] = link_portal[i];
    }
  }
  if (d.node_link_begin) {
    node_link_begin = copy_arr(d.node_link_begin, d.n_polys + 1);
    node_links = copy_arr(d.node_links, d.n_links);
  } else {
    node_link_begin.assign(d.n_polys + 1, 0);
    node_links.clear();
  }
  modified.clear();
  for (uint32_t i = 0; i < d.n_modified; i++) {
    ModifiedNode m;
    for (uint32_t e = d.modified_edge_begin[i]; e < d.modified_edge_begin[i + 1]; e++)
      m.new_boundary.push_back({d.modified_edges[2 * e], d.modified_edges[2 * e + 1]});
    for (uint32_t v = d.modified_vert_begin[i]; v < d.modified_vert_begin[i + 1]; v++)
      m.new_vertices.push_back({d.modified_verts[2 * v], d.modified_verts[2 * v + 1]});
    modified[d.modified_node[i]] = std::move(m);
  }
  node_region_number = copy_arr(d.node_region_number, d.node_region_number ? d.n_polys : 0);
  region_root = copy_arr(d.region_root, d.n_region_numbers);
  possible_links.clear();
  for (uint32_t i = 0; i < d.n_possible_links; i++)
    possible_links[d.possible_link_src[i]].push_back({d.possible_link_kind[i], d.possible_link_dst[i]});
}

// ValidPolygon::get_edge_indices (nav_mesh.rs:945-950): (left, right) =
// (vertices[edge+1 wrapped], vertices[edge]) as GLOBAL vertex indices.
void NavData::edge_indices(uint32_t node, uint32_t edge, uint32_t& left, uint32_t& right) const {
  uint32_t b = poly_edge_begin[node], n = poly_edge_begin[node + 1] - b;
  left = edge_vertex[b + (edge == n - 1 ? 0 : edge + 1)];
  right = edge_vertex[b + edge];
}

// ---------------------------------------------------------------------------
// Sampling — nav_mesh.rs:614-737 (mesh level) and nav_data.rs:269-324.
// ---------------------------------------------------------------------------
static V3 project_to_triangle(V3 t0, V3 t1, V3 t2, V3 point) {
  // nav_mesh.rs:634-665
  V3 d0 = t1 - t0, d1 = t2 - t1, d2 = t0 - t2;
  V2 f0 = xy(d0), f1 = xy(d1), f2 = xy(d2);
  if (perp_dot(f0, xy(point) - xy(t0)) < 0.0f) {
    float s = dot(f0, xy(point) - xy(t0)) / length_squared(f0);
    return d0 * clampf(s, 0.0f, 1.0f) + t0;
  }
  if (perp_dot(f1, xy(point) - xy(t1)) < 0.0f) {
    float s = dot(f1, xy(point) - xy(t1)) / length_squared(f1);
    return d1 * clampf(s, 0.0f, 1.0f) + t1;
  }
  if (perp_dot(f2, xy(point) - xy(t2)) < 0.0f) {
    float s = dot(f2, xy(point) - xy(t2)) / length_squared(f2);
    return d2 * clampf(s, 0.0f, 1.0f) + t2;
  }
  V3 normal = -normalize(cross(d0, d2));
  float height = dot(normal, point - t0) / normal.z;
  return {point.x, point.y, point.z - height};
}

// ValidNavigationMesh::sample_point on one island; `point` is island-local.
// Returns false if nothing is in range.  Linear scan over all polygons, as the
// reference does (nav_mesh.rs:669).
bool NavData::sample_point_island(uint32_t island, V3 point, const lmb_options& o, V3& out_point,
                                  uint32_t& out_node) const {
  const Island& is = islands[island];
  // nav_mesh.rs:619-632 sample box
  V3 smin = point + V3{-o.horizontal_distance, -o.horizontal_distance, -o.distance_below};
  V3 smax = point + V3{o.horizontal_distance, o.horizontal_distance, o.distance_above};
  bool have = false;
  float best = 0.0f;
  for (uint32_t p = is.poly_begin; p < is.poly_end; p++) {
    const float* b = &poly_bounds[(size_t)p * 6];
    // BoundingBox::intersects_bounds (util.rs:201-217); Empty polygon bounds never intersect.
    if (!(b[0] <= b[3])) continue;
    if (!(smin.x <= b[3] && b[0] <= smax.x && smin.y <= b[4] && b[1] <= smax.y && smin.z <= b[5] &&
          b[2] <= smax.z))
      continue;
    auto test_triangle = [&](V3 t0, V3 t1, V3 t2) {
      V3 q = project_to_triangle(t0, t1, t2, point);
      float dh = distance(xy(point), xy(q));
      float dv = q.z - point.z;
      if (dh <= o.horizontal_distance && (-o.distance_below <= dv && dv <= o.distance_above)) {
        float d = dh * o.vertical_preference_ratio + std::fabs(dv);
        if (!have || d < best) {
          have = true;
          best = d;
          out_point = q;
          out_node = p;
        }
      }
    };
    if (is.has_height) {
      const uint32_t* hp = &hpoly[(size_t)p * 4];
      for (uint32_t t = hp[2]; t < hp[2] + hp[3]; t++) {
        const uint8_t* tri = &htris[(size_t)t * 3];
        test_triangle(hverts[hp[0] + tri[0]], hverts[hp[0] + tri[1]], hverts[hp[0] + tri[2]]);
      }
    } else {
      uint32_t eb = poly_edge_begin[p], n = poly_edge_begin[p + 1] - eb;
      for (uint32_t i = 2; i < n; i++)
        test_triangle(vertices[edge_vertex[eb]], vertices[edge_vertex[eb + i - 1]],
                      vertices[edge_vertex[eb + i]]);
    }
  }
  return have;
}

// NavigationData::sample_point (nav_data.rs:269-324).
bool NavData::sample_point(V3 point, const lmb_options& o, V3& out_point, uint32_t& out_node) const {
  bool have = false;
  float best_distance = 0.0f;
  for (uint32_t i = 0; i < n_islands; i++) {
    const Island& is = islands[i];
    V3 rel = is.xf.apply_inverse(point);
    if (is.bounds_empty) continue;
    // add_to_corners (util.rs:146-163) with above/below flipped (nav_data.rs:281-294)
    V3 mn = is.bmin + V3{-o.horizontal_distance, -o.horizontal_distance, -o.distance_above};
    V3 mx = is.bmax + V3{o.horizontal_distance, o.horizontal_distance, o.distance_below};
    V3 size = mx - mn;
    if (!(size.x >= 0.0f && size.y >= 0.0f && size.z >= 0.0f)) continue;  // became Empty
    if (!(mn.x <= rel.x && rel.x <= mx.x && mn.y <= rel.y && rel.y <= mx.y && mn.z <= rel.z &&
          rel.z <= mx.z))
      continue;
    V3 sp;
    uint32_t sn;
    if (!sample_point_island(i, rel, o, sp, sn)) continue;
    float d = distance_squared(rel, sp);
    if (have && d >= best_distance) continue;
    have = true;
    best_distance = d;
    out_point = is.xf.apply(sp);
    out_node = sn;
  }
  return have;
}

// ValidNavigationMesh::sample_point_on_node (nav_mesh.rs:742-814); island-local point.
bool NavData::sample_point_on_node(V3 point, uint32_t node, V3& out) const {
  const Island& is = islands[poly_island[node]];
  auto proj = [&](V3 t0, V3 t1, V3 t2, V3& r) {
    V3 d0 = t1 - t0, d1 = t2 - t1, d2 = t0 - t2;
    const float EPS = -1e-5f;
    if (perp_dot(xy(d0), xy(point) - xy(t0)) < EPS) return false;
    if (perp_dot(xy(d1), xy(point) - xy(t1)) < EPS) return false;
    if (perp_dot(xy(d2), xy(point) - xy(t2)) < EPS) return false;
    V3 normal = -normalize(cross(d0, d2));
    float height = dot(normal, point - t0) / normal.z;
    r = {point.x, point.y, point.z - height};
    return true;
  };
  if (is.has_height) {
    const uint32_t* hp = &hpoly[(size_t)node * 4];
    for (uint32_t t = hp[2]; t < hp[2] + hp[3]; t++) {
      const uint8_t* tri = &htris[(size_t)t * 3];
      if (proj(hverts[hp[0] + tri[0]], hverts[hp[0] + tri[1]], hverts[hp[0] + tri[2]], out))
        return true;
    }
  } else {
    uint32_t eb = poly_edge_begin[node], n = poly_edge_begin[node + 1] - eb;
    for (uint32_t i = 2; i < n; i++)
      if (proj(vertices[edge_vertex[eb]], vertices[edge_vertex[eb + i - 1]],
               vertices[edge_vertex[eb + i]], out))
        return true;
  }
  return false;  // the reference panics here (nav_mesh.rs:811-813)
}

// NavigationData::are_nodes_connected (nav_data.rs:1022-1087).
bool NavData::are_nodes_connected(uint32_t n1, uint32_t n2, const Permitted& permitted) const {
  if (poly_island[n1] == poly_island[n2] && poly_region[n1] == poly_region[n2]) return true;
  if (node_region_number.empty()) return false;
  uint32_t r1 = node_region_number[n1], r2 = node_region_number[n2];
  if (r1 == LMB_NONE || r2 == LMB_NONE) return false;
  if (region_root[r1] == region_root[r2]) return true;
  std::unordered_set<uint32_t> seen;
  std::vector<uint32_t> queue;
  seen.insert(r1);
  queue.push_back(r1);
  for (size_t head = 0; head < queue.size(); head++) {
    uint32_t r = queue[head];
    if (r == r2) return true;
    auto it = possible_links.find(r);
    if (it == possible_links.end()) continue;
    for (auto& pl : it->second) {
      if (!permitted.is_permitted(pl.kind)) continue;
      if (!seen.insert(pl.dst).second) continue;
      queue.push_back(pl.dst);
    }
  }
  return false;
}

// ---------------------------------------------------------------------------
// A* — astar.rs:126-206 with a model of Rust std's BinaryHeap<Reverse<NodeRef>>.
// ---------------------------------------------------------------------------
// NodeRef::partial_cmp (astar.rs:67-77) folded into `a <= b`.
static inline bool noderef_le(const HeapEntry& a, const HeapEntry& b) {
  if (a.estimate < b.estimate) return true;
  if (a.estimate > b.estimate) return false;
  if (a.estimate == b.estimate) {
    // Reverse(a.cost).partial_cmp(&Reverse(b.estimate)) == b.estimate.partial_cmp(&a.cost)
    return b.estimate <= a.cost;  // Less or Equal
  }
  return false;  // NaN: partial_cmp is None
}
// Reverse<T>::le(x, y) = y.0 <= x.0
static inline bool rev_le(const HeapEntry& x, const HeapEntry& y) { return noderef_le(y, x); }

void RustHeap::sift_up(size_t start, size_t pos) {
  HeapEntry hole = data[pos];
  while (pos > start) {
    size_t parent = (pos - 1) / 2;
    if (rev_le(hole, data[parent])) break;
    data[pos] = data[parent];
    pos = parent;
  }
  data[pos] = hole;
}
void RustHeap::push(const HeapEntry& e) {
  size_t old_len = data.size();
  data.push_back(e);
  sift_up(0, old_len);
}
bool RustHeap::pop(HeapEntry& out) {
  if (data.empty()) return false;
  HeapEntry item = data.back();
  data.pop_back();
  if (!data.empty()) {
    std::swap(item, data[0]);
    // sift_down_to_bottom(0)
    size_t end = data.size(), pos = 0;
    HeapEntry hole = data[0];
    size_t child = 1;
    while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
      if (rev_le(data[child], data[child + 1])) child += 1;
      data[pos] = data[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      data[pos] = data[child];
      pos = child;
    }
    data[pos] = hole;
    sift_up(0, pos);
  }
  out = item;
  return true;
}

template <class Problem>
static AStarResult astar_find_path(const Problem& problem) {
  AStarResult res;
  std::unordered_map<uint64_t, float> best_estimates;
  struct Node {
    float cost;
    uint64_t state;
    int64_t prev;
    int64_t action;
  };
  std::vector<Node> all_nodes;
  RustHeap open;
  auto try_add = [&](Node node) {
    float estimate = node.cost + problem.heuristic(node.state);
    auto it = best_estimates.emplace(node.state, std::numeric_limits<float>::infinity()).first;
    if (it->second <= estimate) return;
    it->second = estimate;
    open.push({node.cost, estimate, (uint32_t)all_nodes.size()});
    all_nodes.push_back(node);
  };
  try_add({0.0f, problem.initial_state(), -1, 0});
  HeapEntry cur;
  std::vector<Successor> succ;
  while (open.pop(cur)) {
    const Node current = all_nodes[cur.index];
    if (best_estimates[current.state] < cur.estimate) continue;
    res.explored_nodes += 1;
    if (problem.is_goal(current.state)) {
      res.found = true;
      int64_t i = cur.index;
      while (all_nodes[i].prev >= 0) {
        res.actions.push_back(all_nodes[i].action);
        i = all_nodes[i].prev;
      }
      std::reverse(res.actions.begin(), res.actions.end());
      return res;
    }
    succ.clear();
    problem.successors(current.state, succ);
    for (auto& s : succ) try_add({current.cost + s.cost, s.state, (int64_t)cur.index, s.action});
  }
  return res;
}

// adjacency-list problem of astar_test.rs
struct AdjProblem {
  uint32_t n;
  const uint32_t* begin;
  const float* cost;
  const int32_t* action;
  const uint32_t* state;
  const float* h;
  uint32_t start, end;
  uint64_t initial_state() const { return start; }
  bool is_goal(uint64_t s) const { return s == end; }
  float heuristic(uint64_t s) const { return h[s]; }
  void successors(uint64_t s, std::vector<Successor>& out) const {
    for (uint32_t i = begin[s]; i < begin[s + 1]; i++) out.push_back({cost[i], action[i], state[i]});
  }
};

// ---------------------------------------------------------------------------
// ArchipelagoPathProblem — pathfinding.rs:19-262
// ---------------------------------------------------------------------------
namespace {
constexpr uint64_t ST_START = 0, ST_END = 1;
inline uint64_t st_node_edge(uint32_t node, uint32_t edge) {
  return (2ull << 60) | ((uint64_t)node << 16) | edge;
}
inline uint64_t st_link(uint32_t link) { return (3ull << 60) | link; }
inline int st_kind(uint64_t s) { return s < 2 ? (int)s : (int)(s >> 60); }
constexpr int64_t ACT_GO_TO_END = -1;
inline int64_t act_edge(uint32_t e) { return e; }
inline int64_t act_link(uint32_t l) { return (int64_t)LMB_PORTAL_LINK | l; }

struct PathProblem {
  const NavData* nav;
  uint32_t start_node, end_node;
  V3 start_point, end_point;
  float cheapest;
  const std::unordered_map<uint32_t, float>* overrides;
  const Permitted* permitted;

  float type_cost(uint32_t type) const {  // pathfinding.rs:73-77
    auto it = overrides->find(type);
    if (it != overrides->end()) return it->second;
    auto jt = nav->type_cost.find(type);
    return jt != nav->type_cost.end() ? jt->second : 1.0f;
  }
  V3 edge_midpoint_world(uint32_t node, uint32_t edge) const {
    uint32_t i, j;
    nav->edge_indices(node, edge, i, j);
    return nav->islands[nav->poly_island[node]].xf.apply(midpoint(nav->vertices[i], nav->vertices[j]));
  }
  uint64_t initial_state() const { return ST_START; }
  bool is_goal(uint64_t s) const { return s == ST_END; }
  float heuristic(uint64_t s) const {  // pathfinding.rs:233-257
    V3 wp;
    switch (st_kind(s)) {
      case 0: wp = start_point; break;
      case 1: return 0.0f;
      case 2: wp = edge_midpoint_world((uint32_t)((s >> 16) & 0xFFFFFFFFu), (uint32_t)(s & 0xFFFF)); break;
      default: {
        uint32_t l = (uint32_t)(s & 0xFFFFFFFFu);
        auto portal = nav->link_kind[l] == LMB_LINK_BOUNDARY ? nav->link_portal[l] : nav->link_dst_portal[l];
        wp = midpoint(portal.first, portal.second);
      }
    }
    return distance(wp, end_point) * cheapest;
  }
  void successors(uint64_t s, std::vector<Successor>& out) const {  // pathfinding.rs:89-231
    uint32_t node;
    V3 point;
    int64_t ignore_edge = -1, ignore_link = -1;
    switch (st_kind(s)) {
      case 0:
        node = start_node;
        point = start_point;
        break;
      case 2: {
        node = (uint32_t)((s >> 16) & 0xFFFFFFFFu);
        uint32_t edge = (uint32_t)(s & 0xFFFF);
        point = edge_midpoint_world(node, edge);
        ignore_edge = edge;
        break;
      }
      case 3: {
        uint32_t l = (uint32_t)(s & 0xFFFFFFFFu);
        node = nav->link_dst_node[l];
        std::pair<V3, V3> portal;
        if (nav->link_kind[l] == LMB_LINK_BOUNDARY) {
          portal = nav->link_portal[l];
          ignore_link = nav->link_reverse[l];
        } else {
          portal = nav->link_dst_portal[l];
        }
        point = midpoint(portal.first, portal.second);
        break;
      }
      default: return;  // unreachable!() in the reference
    }
    float current_cost = type_cost(nav->poly_type[node]);
    if (node == end_node) {
      out.push_back({distance(point, end_point) * current_cost, ACT_GO_TO_END, ST_END});
      return;
    }
    uint32_t eb = nav->poly_edge_begin[node], n = nav->poly_edge_begin[node + 1] - eb;
    for (uint32_t e = 0; e < n; e++) {
      uint32_t nb = nav->edge_neighbor[eb + e];
      if (nb == LMB_NONE) continue;
      if ((int64_t)e == ignore_edge) continue;
      float target_cost = type_cost(nav->poly_type[nb]);
      if (!std::isfinite(target_cost)) continue;
      float cost = distance(point, edge_midpoint_world(node, e)) * current_cost;
      out.push_back({cost, act_edge(e), st_node_edge(nb, nav->edge_reverse[eb + e])});
    }
    // off-mesh links in canonical ascending order (the reference iterates a HashSet)
    for (uint32_t k = nav->node_link_begin[node]; k < nav->node_link_begin[node + 1]; k++) {
      uint32_t l = nav->node_links[k];
      if ((int64_t)l == ignore_link) continue;
      float dst_cost = type_cost(nav->link_dst_type[l]);
      if (!std::isfinite(dst_cost)) continue;
      float link_cost = 0.0f;
      if (nav->link_kind[l] == LMB_LINK_ANIMATION) {
        if (!permitted->is_permitted(nav->link_anim_kind[l])) continue;
        link_cost = nav->link_cost[l];
      }
      float cost = distance(point, midpoint(nav->link_portal[l].first, nav->link_portal[l].second)) *
                       current_cost +
                   link_cost;
      out.push_back({cost, act_link(l), st_link(l)});
    }
  }
};
}  // namespace

// pathfinding::find_path (pathfinding.rs:277-382)
PathResult find_path(const NavData& nav, uint32_t start_node, V3 start_point, uint32_t end_node,
                     V3 end_point, const std::unordered_map<uint32_t, float>& overrides,
                     const Permitted& permitted) {
  PathResult out;
  if (!nav.are_nodes_connected(start_node, end_node, permitted)) return out;
  PathProblem problem;
  problem.nav = &nav;
  problem.start_node = start_node;
  problem.end_node = end_node;
  problem.start_point = start_point;
  problem.end_point = end_point;
  problem.overrides = &overrides;
  problem.permitted = &permitted;
  // cheapest_type_index_cost (pathfinding.rs:300-314): FloatOrd min, NaN largest
  float cheapest = 1.0f;
  for (auto& kv : nav.type_cost) {
    auto it = overrides.find(kv.first);
    float c = it != overrides.end() ? it->second : kv.second;
    if (!std::isfinite(c)) continue;
    if (c < cheapest) cheapest = c;
  }
  problem.cheapest = cheapest;
  AStarResult r = astar_find_path(problem);
  out.explored_nodes = r.explored_nodes;
  if (!r.found) return out;
  Path path;
  path.start_point = start_point;
  path.end_point = end_point;
  path.island_segments.push_back({nav.poly_island[start_node], {start_node}, {}});
  for (int64_t step : r.actions) {
    IslandSegment& last = path.island_segments.back();
    uint32_t previous_node = last.corridor.back();
    if (step == ACT_GO_TO_END) {
    } else if (step & (int64_t)LMB_PORTAL_LINK) {
      uint32_t l = (uint32_t)(step & 0x7FFFFFFF);
      path.off_mesh_link_segments.push_back({previous_node, nav.link_dst_node[l], l});
      path.island_segments.push_back({nav.poly_island[nav.link_dst_node[l]], {nav.link_dst_node[l]}, {}});
    } else {
      uint32_t e = (uint32_t)step;
      last.corridor.push_back(nav.edge_neighbor[nav.poly_edge_begin[previous_node] + e]);
      last.portal_edge_index.push_back(e);
    }
  }
  out.path = std::move(path);
  return out;
}

// ---------------------------------------------------------------------------
// Path — path.rs
// ---------------------------------------------------------------------------
PathIndex PathIndex::next(const Path& path) const {  // path.rs:153-176
  uint32_t np = portal_index + 1;
  size_t len = path.island_segments[segment_index].portal_edge_index.size();
  if (np <= len) return {segment_index, np};
  if ((size_t)segment_index + 1 < path.island_segments.size()) return {segment_index + 1, 0};
  return {segment_index, np};
}

PathIndex Path::last_index() const {
  uint32_t s = (uint32_t)island_segments.size() - 1;
  return {s, (uint32_t)island_segments[s].portal_edge_index.size()};
}

std::optional<PathIndex> Path::find_index_of_node(const NavData& nav, uint32_t node) const {
  for (uint32_t s = 0; s < island_segments.size(); s++) {
    if (nav.poly_island[node] != island_segments[s].island) continue;
    for (uint32_t p = 0; p < island_segments[s].corridor.size(); p++)
      if (island_segments[s].corridor[p] == node) return PathIndex{s, p};
  }
  return std::nullopt;
}
std::optional<PathIndex> Path::find_index_of_node_rev(const NavData& nav, uint32_t node) const {
  for (uint32_t s = (uint32_t)island_segments.size(); s-- > 0;) {
    if (nav.poly_island[node] != island_segments[s].island) continue;
    for (uint32_t p = (uint32_t)island_segments[s].corridor.size(); p-- > 0;)
      if (island_segments[s].corridor[p] == node) return PathIndex{s, p};
  }
  return std::nullopt;
}

// flatten to (node, portal code) pairs in PathIndex order
void Path::flatten(std::vector<uint32_t>& nodes, std::vector<uint32_t>& portals) const {
  nodes.clear();
  portals.clear();
  for (size_t s = 0; s < island_segments.size(); s++) {
    const IslandSegment& seg = island_segments[s];
    for (size_t p = 0; p < seg.corridor.size(); p++) {
      nodes.push_back(seg.corridor[p]);
      if (p < seg.portal_edge_index.size())
        portals.push_back(seg.portal_edge_index[p]);
      else if (s + 1 < island_segments.size())
        portals.push_back(LMB_PORTAL_LINK | off_mesh_link_segments[s].link);
      else
        portals.push_back(LMB_NONE);
    }
  }
}
Path Path::from_flat(const NavData& nav, uint32_t len, const uint32_t* nodes, const uint32_t* portals) {
  Path path;
  path.start_point = path.end_point = {0, 0, 0};
  bool new_seg = true;
  for (uint32_t i = 0; i < len; i++) {
    if (new_seg) {
      path.island_segments.push_back({nav.poly_island[nodes[i]], {}, {}});
      new_seg = false;
    }
    IslandSegment& seg = path.island_segments.back();
    seg.corridor.push_back(nodes[i]);
    if (portals[i] == LMB_NONE) {
    } else if (portals[i] & LMB_PORTAL_LINK) {
      uint32_t l = portals[i] & 0x7FFFFFFF;
      path.off_mesh_link_segments.push_back({nodes[i], nav.link_dst_node[l], l});
      new_seg = true;
    } else {
      seg.portal_edge_index.push_back(portals[i]);
    }
  }
  return path;
}
uint32_t Path::flat_index(PathIndex idx) const {
  uint32_t base = 0;
  for (uint32_t s = 0; s < idx.segment_index; s++) base += (uint32_t)island_segments[s].corridor.size();
  return base + idx.portal_index;
}
PathIndex Path::from_flat_index(uint32_t flat) const {
  uint32_t s = 0;
  while (s + 1 < island_segments.size() && flat >= island_segments[s].corridor.size()) {
    flat -= (uint32_t)island_segments[s].corridor.size();
    s++;
  }
  return {s, flat};
}

namespace {
struct Portal {
  bool walkable;
  V3 a, b;  // Walkable(left,right) or AnimationLink start_portal
  V3 end_a, end_b;
  uint32_t link_id, start_node, end_node;
};
// Path::get_portal_endpoints (path.rs:81-129, 204-218)
Portal get_portal_endpoints(const Path& path, const NavData& nav, PathIndex idx) {
  Portal p{};
  const IslandSegment& seg = path.island_segments[idx.segment_index];
  if (idx.portal_index == seg.portal_edge_index.size()) {
    const OffMeshLinkSegment& ls = path.off_mesh_link_segments[idx.segment_index];
    auto portal = nav.link_portal[ls.link];
    if (nav.link_kind[ls.link] == LMB_LINK_BOUNDARY) {
      p.walkable = true;
      p.a = portal.first;
      p.b = portal.second;
    } else {
      p.walkable = false;
      p.a = portal.first;
      p.b = portal.second;
      p.end_a = nav.link_dst_portal[ls.link].first;
      p.end_b = nav.link_dst_portal[ls.link].second;
      p.link_id = nav.link_anim_id[ls.link];
      p.start_node = ls.starting_node;
      p.end_node = ls.end_node;
    }
  } else {
    uint32_t node = seg.corridor[idx.portal_index];
    uint32_t edge = seg.portal_edge_index[idx.portal_index];
    uint32_t l, r;
    nav.edge_indices(node, edge, l, r);
    const Xform& xf = nav.islands[seg.island].xf;
    p.walkable = true;
    p.a = xf.apply(nav.vertices[l]);
    p.b = xf.apply(nav.vertices[r]);
  }
  return p;
}
// geometry.rs:176-189
std::pair<V3, float> project_point_to_line_segment(V3 point, V3 start, V3 end) {
  V3 dir = end - start;
  float l2 = length_squared(dir);
  if (l2 == 0.0f) return {start, 0.0f};
  V3 rel = point - start;
  float fraction = clampf(dot(rel, dir) / l2, 0.0f, 1.0f);
  return {dir * fraction + start, fraction};
}
inline float triangle_area_2(V3 p0, V3 p1, V3 p2) {
  return perp_dot(xy(p1) - xy(p0), xy(p2) - xy(p0));
}
}  // namespace

// Path::find_next_point_in_straight_path (path.rs:227-368)
std::pair<PathIndex, StraightPathStep> Path::find_next_point_in_straight_path(
    const NavData& nav, PathIndex start_index, V3 start_point, PathIndex end_index,
    V3 end_point) const {
  V3 apex = start_point;
  PathIndex left_index = start_index, right_index = start_index;
  bool end_is_link = false;
  StraightPathStep end_link{};
  V3 current_left, current_right;
  auto make_link_step = [&](const Portal& p) {
    auto pr = project_point_to_line_segment(apex, p.a, p.b);
    StraightPathStep st{};
    st.is_link = true;
    st.point = pr.first;
    st.end_point = lerp(p.end_a, p.end_b, pr.second);
    st.link_id = p.link_id;
    st.start_node = p.start_node;
    st.end_node = p.end_node;
    return st;
  };
  if (start_index == end_index) {
    current_left = current_right = end_point;
  } else {
    Portal p = get_portal_endpoints(*this, nav, start_index);
    if (p.walkable) {
      current_left = p.a;
      current_right = p.b;
    } else {
      return {start_index.next(*this), make_link_step(p)};
    }
  }
  PathIndex portal_index = start_index.next(*this);
  while (portal_index <= end_index) {
    V3 portal_left, portal_right;
    if (portal_index == end_index) {
      portal_left = portal_right = end_point;
    } else {
      Portal p = get_portal_endpoints(*this, nav, portal_index);
      if (p.walkable) {
        portal_left = p.a;
        portal_right = p.b;
      } else {
        end_link = make_link_step(p);
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
        return {left_index, StraightPathStep{false, current_left}};
      }
    }
    if (triangle_area_2(apex, current_left, portal_left) <= 0.0f) {
      if (triangle_area_2(apex, current_right, portal_left) >= 0.0f) {
        left_index = portal_index;
        current_left = portal_left;
      } else {
        return {right_index, StraightPathStep{false, current_right}};
      }
    }
    portal_index = portal_index.next(*this);
  }
  if (!end_is_link) return {end_index, StraightPathStep{false, end_point}};
  return {portal_index, end_link};
}

// ---------------------------------------------------------------------------
// Agent logic — agent.rs:312-475
// ---------------------------------------------------------------------------
static float reach_distance_or_radius(const AgentIn& a) {
  return (a.reach_distance >= 0.0f) ? a.reach_distance : a.radius;  // NaN -> radius
}
static float animation_link_reached_distance(const AgentIn& a) {  // agent.rs:312-323
  if (a.anim_link_reach_distance >= 0.0f) return a.anim_link_reach_distance;
  return reach_distance_or_radius(a);
}

// Agent::has_reached_target (agent.rs:330-404)
static bool has_reached_target(const AgentIn& a, const Path& path, const NavData& nav,
                               V3 sampled_point, std::pair<PathIndex, StraightPathStep> next_waypoint,
                               std::pair<PathIndex, V3> target_waypoint) {
  float distance_ = reach_distance_or_radius(a);
  switch (a.reach_condition) {
    case LMB_REACH_DISTANCE:
      return distance_squared(sampled_point, target_waypoint.second) < distance_ * distance_;
    case LMB_REACH_VISIBLE_AT_DISTANCE:
      if (next_waypoint.second.is_link) return false;
      return next_waypoint.first == target_waypoint.first &&
             distance_squared(sampled_point, next_waypoint.second.point) < distance_ * distance_;
    default: {
      if (distance_squared(sampled_point, target_waypoint.second) > distance_ * distance_) return false;
      if (next_waypoint.first == target_waypoint.first) return true;
      if (next_waypoint.second.is_link) return false;
      float straight = distance(sampled_point, next_waypoint.second.point);
      std::pair<PathIndex, V3> current{next_waypoint.first, next_waypoint.second.point};
      while (!(current.first == target_waypoint.first) && straight < distance_) {
        auto nw = path.find_next_point_in_straight_path(nav, current.first, current.second,
                                                        target_waypoint.first, target_waypoint.second);
        if (nw.second.is_link) return false;
        straight += distance(current.second, nw.second.point);
        current = {nw.first, nw.second.point};
      }
      return straight < distance_;
    }
  }
}

// ---------------------------------------------------------------------------
// Archipelago::update — lib.rs:270-534
// ---------------------------------------------------------------------------
void Ctx::step_begin(const lmb_agents_in& in, const lmb_characters_in* chars, const lmb_options& opt,
                     float dt, uint32_t n_total, uint32_t first) {
  pathing_results.clear();
  const uint32_t n = in.n;
  f_agents.assign(n, AgentIn{});
  std::vector<AgentIn>& ag = f_agents;
  for (uint32_t i = 0; i < n; i++) {
    AgentIn& a = ag[i];
    a.slot = in.slot[i];
    a.flags = in.flags[i];
    a.position = {in.position[3 * i], in.position[3 * i + 1], in.position[3 * i + 2]};
    a.velocity = {in.velocity[3 * i], in.velocity[3 * i + 1], in.velocity[3 * i + 2]};
    if (in.target) a.target = {in.target[3 * i], in.target[3 * i + 1], in.target[3 * i + 2]};
    a.radius = in.radius[i];
    a.desired_speed = in.desired_speed[i];
    a.max_speed = in.max_speed[i];
    a.reach_condition = in.reach_condition ? in.reach_condition[i] : LMB_REACH_DISTANCE;
    a.reach_distance = in.reach_distance ? in.reach_distance[i] : -1.0f;
    a.anim_link_reach_distance = in.anim_link_reach_distance ? in.anim_link_reach_distance[i] : -1.0f;
    if (in.override_begin)
      for (uint32_t k = in.override_begin[i]; k < in.override_begin[i + 1]; k++)
        a.overrides[in.override_type[k]] = in.override_cost[k];
    a.permitted.all = (a.flags & LMB_AGENT_PERMIT_ALL_LINKS) != 0;
    if (!a.permitted.all && in.permitted_begin)
      for (uint32_t k = in.permitted_begin[i]; k < in.permitted_begin[i + 1]; k++)
        a.permitted.kinds.insert(in.permitted_kind[k]);
    if (a.flags & LMB_AGENT_RESET) agent_state.erase(a.slot);
  }
  auto st = [&](uint32_t i) -> AgentPersist& { return agent_state[ag[i].slot]; };

  // PHASE A: sample agents + targets (lib.rs:286-327)
  f_agent_node.assign(n, Sampled{});
  std::vector<Sampled>& agent_node = f_agent_node;
  std::vector<Sampled> target_node(n);
  for (uint32_t i = 0; i < n; i++) {
    AgentPersist& s = st(i);
    if (ag[i].flags & LMB_AGENT_PAUSED) {
      s.state = LMB_STATE_PAUSED;
      continue;
    }
    if (ag[i].flags & LMB_AGENT_USING_ANIMATION_LINK) {
      s.state = LMB_STATE_USING_ANIMATION_LINK;
      continue;
    }
    agent_node[i].valid = nav.sample_point(ag[i].position, opt, agent_node[i].point, agent_node[i].node);
    if (!agent_node[i].valid) continue;
    if (ag[i].flags & LMB_AGENT_HAS_TARGET)
      target_node[i].valid = nav.sample_point(ag[i].target, opt, target_node[i].point, target_node[i].node);
  }
  // PHASE A': characters (lib.rs:329-341)
  f_chars.assign(chars ? chars->n : 0, CharacterIn{});
  std::vector<CharacterIn>& ch = f_chars;
  for (uint32_t i = 0; i < ch.size(); i++) {
    ch[i].position = {chars->position[3 * i], chars->position[3 * i + 1], chars->position[3 * i + 2]};
    ch[i].velocity = {chars->velocity[3 * i], chars->velocity[3 * i + 1], chars->velocity[3 * i + 2]};
    ch[i].radius = chars->radius[i];
    uint32_t node;
    ch[i].valid = nav.sample_point(ch[i].position, opt, ch[i].point, node);
  }

  // PHASE B: repath decision + A* (lib.rs:343-424).  `path_invalid` is this frame's
  // `!path.is_valid(..)`, computed by orc_navdata_update.
  std::vector<std::pair<PathIndex, PathIndex>> follow(n);
  std::vector<bool> has_follow(n, false);
  for (uint32_t i = 0; i < n; i++) {
    AgentPersist& s = st(i);
    s.reached_link_id = LMB_NONE;  // current_animation_link = None (lib.rs:348)
    const bool path_invalid = s.path_invalid;
    s.path_invalid = false;
    if (ag[i].flags & (LMB_AGENT_PAUSED | LMB_AGENT_USING_ANIMATION_LINK)) {
      if (s.path && path_invalid) s.path.reset();  // lib.rs:350-358
      continue;
    }
    // does_agent_need_repath (agent.rs:425-475)
    if (!(ag[i].flags & LMB_AGENT_HAS_TARGET)) {
      if (s.path) {
        s.state = LMB_STATE_IDLE;
        s.path.reset();
      }
      continue;
    }
    if (!agent_node[i].valid) {
      s.state = LMB_STATE_AGENT_NOT_ON_NAV_MESH;
      s.path.reset();
      continue;
    }
    if (!target_node[i].valid) {
      s.state = LMB_STATE_TARGET_NOT_ON_NAV_MESH;
      s.path.reset();
      continue;
    }
    bool repath = true;
    if (s.path && !path_invalid) {  // agent.rs:449-456
      auto ai = s.path->find_index_of_node(nav, agent_node[i].node);
      if (ai) {
        auto ti = s.path->find_index_of_node_rev(nav, target_node[i].node);
        if (ti && !(*ai > *ti)) {
          follow[i] = {*ai, *ti};
          has_follow[i] = true;
          repath = false;
        }
      }
    }
    if (!repath) continue;
    s.path.reset();
    PathResult pr = find_path(nav, agent_node[i].node, agent_node[i].point, target_node[i].node,
                              target_node[i].point, ag[i].overrides, ag[i].permitted);
    pathing_results.push_back({i, pr.path ? 1u : 0u, pr.explored_nodes});
    if (!pr.path) {
      s.state = LMB_STATE_NO_PATH;
      continue;
    }
    follow[i] = {PathIndex{0, 0}, pr.path->last_index()};
    has_follow[i] = true;
    s.path = std::move(pr.path);
  }

  // PHASE C: follow path (lib.rs:426-523)
  f_waypoints.assign(n, Waypoint{});
  for (uint32_t i = 0; i < n; i++) {
    AgentPersist& s = st(i);
    if (!s.path) {
      s.desired_move = {0, 0, 0};
      continue;
    }
    if (!agent_node[i].valid) continue;  // paused with a path
    V3 agent_point = agent_node[i].point;
    V3 target_point = target_node[i].point;
    auto next_waypoint = s.path->find_next_point_in_straight_path(nav, follow[i].first, agent_point,
                                                                  follow[i].second, target_point);
    {
      const StraightPathStep& w = next_waypoint.second;
      f_waypoints[i].kind = w.is_link ? 1u : 0u;
      f_waypoints[i].link = w.is_link ? w.link_id : LMB_NONE;
      f_waypoints[i].point = w.point;
      f_waypoints[i].end_point = w.is_link ? w.end_point : w.point;
    }
    if (has_reached_target(ag[i], *s.path, nav, agent_point, next_waypoint,
                           {follow[i].second, target_point})) {
      s.desired_move = {0, 0, 0};
      s.state = LMB_STATE_REACHED_TARGET;
    } else {
      V3 waypoint;
      if (!next_waypoint.second.is_link) {
        s.state = LMB_STATE_MOVING;
        waypoint = next_waypoint.second.point;
      } else {
        auto refine = [&](V3 point, uint32_t node) {
          const Xform& xf = nav.islands[nav.poly_island[node]].xf;
          V3 local = xf.apply_inverse(point), r = local;
          nav.sample_point_on_node(local, node, r);
          return xf.apply(r);
        };
        V3 start_point = refine(next_waypoint.second.point, next_waypoint.second.start_node);
        V3 end_point = refine(next_waypoint.second.end_point, next_waypoint.second.end_node);
        float d = distance(agent_point, start_point);
        if (d <= animation_link_reached_distance(ag[i])) {
          s.state = LMB_STATE_REACHED_ANIMATION_LINK;
          s.reached_link_id = next_waypoint.second.link_id;
          s.reached_link_start = start_point;
          s.reached_link_end = end_point;
        } else {
          s.state = LMB_STATE_MOVING;
        }
        waypoint = start_point;
      }
      V2 dm = normalize_or_zero(xy(waypoint - ag[i].position)) * ag[i].desired_speed;
      s.desired_move = {dm.x, dm.y, 0.0f};
    }
  }

  // PHASE D, first half: this rank's avoidance records into the halo buffer
  halo.assign(n_total, AvoidRec{});
  halo_first = first;
  std::vector<AgentPersist*> persist(n);
  for (uint32_t i = 0; i < n; i++) persist[i] = &st(i);
  build_avoid_records(ag, persist, agent_node, opt, halo.data() + first);
}

void Ctx::step_end(const lmb_options& opt, float dt, const lmb_agents_out& out) {
  std::vector<AgentIn>& ag = f_agents;
  const uint32_t n = (uint32_t)ag.size();
  auto st = [&](uint32_t i) -> AgentPersist& { return agent_state[ag[i].slot]; };
  // PHASE D: avoidance (lib.rs:525-533) against every rank's records
  if (!(flags & LMB_CONFIG_NO_AVOIDANCE)) {
    std::vector<AgentPersist*> persist(n);
    for (uint32_t i = 0; i < n; i++) persist[i] = &st(i);
    f_avoid_debug.assign(n, AvoidDebugData{});
    apply_avoidance_to_agents(nav, halo, halo_first, ag, persist, f_agent_node, f_chars, opt, dt, &f_avoid_debug);
  }

  for (uint32_t i = 0; i < n; i++) {
    AgentPersist& s = st(i);
    if (out.desired_velocity) {
      out.desired_velocity[3 * i] = s.desired_move.x;
      out.desired_velocity[3 * i + 1] = s.desired_move.y;
      out.desired_velocity[3 * i + 2] = s.desired_move.z;
    }
    if (out.state) out.state[i] = s.state;
    if (out.reached_link_id) {
      out.reached_link_id[i] = s.reached_link_id;
      if (out.reached_link_start && out.reached_link_end) {
        V3 a = s.reached_link_id != LMB_NONE ? s.reached_link_start : V3{0, 0, 0};
        V3 b = s.reached_link_id != LMB_NONE ? s.reached_link_end : V3{0, 0, 0};
        out.reached_link_start[3 * i] = a.x;
        out.reached_link_start[3 * i + 1] = a.y;
        out.reached_link_start[3 * i + 2] = a.z;
        out.reached_link_end[3 * i] = b.x;
        out.reached_link_end[3 * i + 1] = b.y;
        out.reached_link_end[3 * i + 2] = b.z;
      }
    }
  }
}

}  // namespace orc

// ---------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------
using namespace orc;

extern "C" {

orc_ctx* orc_create(const lmb_navdata_desc* desc, uint32_t flags) {
  if (!desc || desc->struct_size != sizeof(lmb_navdata_desc)) return nullptr;
  orc_ctx* h = new orc_ctx();
  h->c.flags = flags;
  h->c.nav.load(*desc);
  h->c.has_nav = true;
  return h;
}
void orc_destroy(orc_ctx* h) { delete h; }
int32_t orc_navdata_upload(orc_ctx* h, const lmb_navdata_desc* desc) { return orc_navdata_update(h, desc, nullptr); }

// Path::is_valid (path.rs:381-400): no island segment on an invalidated island, no off-mesh link
// segment over an invalidated link.
bool Path::is_valid(const uint32_t* island_map, const uint32_t* link_map) const {
  for (const IslandSegment& seg : island_segments)
    if (island_map[seg.island] == LMB_NONE) return false;
  for (const OffMeshLinkSegment& seg : off_mesh_link_segments)
    if (link_map[seg.link] == LMB_NONE) return false;
  return true;
}

int32_t orc_navdata_update(orc_ctx* h, const lmb_navdata_desc* desc, const lmb_nav_remap* remap) {
  if (!h || !desc || desc->struct_size != sizeof(lmb_navdata_desc)) return LMB_ERR_INVALID_ARGUMENT;
  NavData& nav = h->c.nav;
  if (!remap || !h->c.has_nav) {
    nav.load(*desc);
    h->c.has_nav = true;
    for (auto& kv : h->c.agent_state) kv.second.path.reset();
    return LMB_OK;
  }
  if (remap->struct_size != sizeof(lmb_nav_remap) || remap->n_old_islands != nav.islands.size() ||
      remap->n_old_links != nav.link_dst_node.size())
    return LMB_ERR_INVALID_ARGUMENT;
  // `invalidated_islands` / `invalidated_off_mesh_links` (nav_data.rs:326-350) are the entries
  // mapped to LMB_NONE.  Valid paths are re-expressed in the new flattening's ids.
  std::vector<uint32_t> old_begin(nav.islands.size());
  for (size_t i = 0; i < nav.islands.size(); i++) old_begin[i] = nav.islands[i].poly_begin;
  for (auto& kv : h->c.agent_state) {
    AgentPersist& s = kv.second;
    if (!s.path) continue;
    if (s.path_invalid || !s.path->is_valid(remap->island_map, remap->link_map)) {
      s.path_invalid = true;
      continue;
    }
    for (IslandSegment& seg : s.path->island_segments) {
      const uint32_t ni = remap->island_map[seg.island];
      for (uint32_t& node : seg.corridor) node = desc->island_poly_begin[ni] + (node - old_begin[seg.island]);
      seg.island = ni;
    }
    for (OffMeshLinkSegment& seg : s.path->off_mesh_link_segments) {
      auto renode = [&](uint32_t node) {
        const uint32_t oi = nav.poly_island[node];
        return desc->island_poly_begin[remap->island_map[oi]] + (node - old_begin[oi]);
      };
      seg.starting_node = renode(seg.starting_node);
      seg.end_node = renode(seg.end_node);
      seg.link = remap->link_map[seg.link];
    }
  }
  nav.load(*desc);
  return LMB_OK;
}

int32_t orc_sample_points(orc_ctx* h, uint32_t n, const float* points, const lmb_options* o,
                          float* out_points, uint32_t* out_nodes) {
  for (uint32_t i = 0; i < n; i++) {
    V3 p{points[3 * i], points[3 * i + 1], points[3 * i + 2]}, q{0, 0, 0};
    uint32_t node = LMB_NONE;
    if (!h->c.nav.sample_point(p, *o, q, node)) {
      node = LMB_NONE;
      q = {0, 0, 0};
    }
    out_points[3 * i] = q.x;
    out_points[3 * i + 1] = q.y;
    out_points[3 * i + 2] = q.z;
    out_nodes[i] = node;
  }
  return LMB_OK;
}

// permitted_animation_links of query i of a batched query (see lmb_find_paths)
static Permitted query_permitted(uint32_t i, const uint32_t* permit_all, const uint32_t* permitted_begin,
                                 const uint32_t* permitted_kind) {
  Permitted p;
  p.all = !permit_all || permit_all[i] != 0;
  if (!p.all && permitted_begin)
    for (uint32_t k = permitted_begin[i]; k < permitted_begin[i + 1]; k++) p.kinds.insert(permitted_kind[k]);
  return p;
}

int32_t orc_find_paths(orc_ctx* h, uint32_t n, const uint32_t* start_node, const float* start_point,
                       const uint32_t* end_node, const float* end_point,
                       const uint32_t* override_begin, const uint32_t* override_type,
                       const float* override_cost, uint32_t* out_success, uint32_t* out_explored,
                       uint32_t* out_path_begin, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t path_capacity, const uint32_t* permit_all, const uint32_t* permitted_begin,
                       const uint32_t* permitted_kind) {
  uint32_t cursor = 0;
  std::vector<uint32_t> nodes, portals;
  int32_t status = LMB_OK;
  for (uint32_t i = 0; i < n; i++) {
    std::unordered_map<uint32_t, float> ov;
    if (override_begin)
      for (uint32_t k = override_begin[i]; k < override_begin[i + 1]; k++)
        ov[override_type[k]] = override_cost[k];
    V3 sp{start_point[3 * i], start_point[3 * i + 1], start_point[3 * i + 2]};
    V3 ep{end_point[3 * i], end_point[3 * i + 1], end_point[3 * i + 2]};
    PathResult pr = find_path(h->c.nav, start_node[i], sp, end_node[i], ep, ov,
                              query_permitted(i, permit_all, permitted_begin, permitted_kind));
    out_success[i] = pr.path ? 1 : 0;
    out_explored[i] = pr.explored_nodes;
    out_path_begin[i] = cursor;
    if (pr.path) {
      pr.path->flatten(nodes, portals);
      if (cursor + nodes.size() > path_capacity) {
        status = LMB_ERR_CAPACITY;
      } else {
        std::copy(nodes.begin(), nodes.end(), out_nodes + cursor);
        std::copy(portals.begin(), portals.end(), out_portals + cursor);
      }
      cursor += (uint32_t)nodes.size();
    }
  }
  out_path_begin[n] = cursor;
  return status;
}

int32_t orc_find_straight_paths(orc_ctx* h, uint32_t n, const uint32_t* start_node,
                                const float* start_point, const uint32_t* end_node,
                                const float* end_point, const uint32_t* override_begin,
                                const uint32_t* override_type, const float* override_cost,
                                uint32_t* out_success, uint32_t* out_step_begin, uint32_t* out_kind,
                                float* out_points, uint32_t* out_link, uint32_t step_capacity,
                                const uint32_t* permit_all, const uint32_t* permitted_begin,
                                const uint32_t* permitted_kind) {
  // query.rs:185-273
  uint32_t cursor = 0;
  int32_t status = LMB_OK;
  auto emit = [&](uint32_t kind, V3 a, V3 b, uint32_t link) {
    if (cursor < step_capacity) {
      out_kind[cursor] = kind;
      float* p = out_points + 6 * (size_t)cursor;
      p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = b.x; p[4] = b.y; p[5] = b.z;
      out_link[cursor] = link;
    } else {
      status = LMB_ERR_CAPACITY;
    }
    cursor++;
  };
  for (uint32_t i = 0; i < n; i++) {
    out_step_begin[i] = cursor;
    std::unordered_map<uint32_t, float> ov;
    if (override_begin)
      for (uint32_t k = override_begin[i]; k < override_begin[i + 1]; k++) {
        if (!(override_cost[k] > 0.0f)) return LMB_ERR_INVALID_ARGUMENT;
        ov[override_type[k]] = override_cost[k];
      }
    V3 sp{start_point[3 * i], start_point[3 * i + 1], start_point[3 * i + 2]};
    V3 ep{end_point[3 * i], end_point[3 * i + 1], end_point[3 * i + 2]};
    PathResult pr = find_path(h->c.nav, start_node[i], sp, end_node[i], ep, ov,
                              query_permitted(i, permit_all, permitted_begin, permitted_kind));
    out_success[i] = pr.path ? 1 : 0;
    if (!pr.path) continue;
    const Path& path = *pr.path;
    PathIndex current_index{0, 0};
    V3 current_point = sp;
    PathIndex last_index = path.last_index();
    emit(0, sp, sp, LMB_NONE);
    if (current_index == last_index) {
      emit(0, ep, ep, LMB_NONE);
      continue;
    }
    bool last_was_link = false;
    while (!(current_index == last_index) || last_was_link) {
      auto r = path.find_next_point_in_straight_path(h->c.nav, current_index, current_point, last_index, ep);
      current_index = r.first;
      if (!r.second.is_link) {
        current_point = r.second.point;
        emit(0, r.second.point, r.second.point, LMB_NONE);
        last_was_link = false;
      } else {
        current_point = r.second.end_point;
        emit(1, r.second.point, r.second.end_point, r.second.link_id);
        last_was_link = true;
      }
    }
  }
  out_step_begin[n] = cursor;
  return status;
}

int32_t orc_step(orc_ctx* h, const lmb_agents_in* agents, const lmb_characters_in* characters,
                 const lmb_options* options, float dt, const lmb_agents_out* out) {
  if (!h || !agents || !options || !out) return LMB_ERR_INVALID_ARGUMENT;
  h->c.step(*agents, characters, *options, dt, *out);
  return LMB_OK;
}
int32_t orc_step_begin(orc_ctx* h, const lmb_agents_in* agents, const lmb_characters_in* characters,
                       const lmb_options* options, float dt, uint32_t n_total, uint32_t first) {
  if (!h || !agents || !options || first + agents->n > n_total) return LMB_ERR_INVALID_ARGUMENT;
  h->c.step_begin(*agents, characters, *options, dt, n_total, first);
  return LMB_OK;
}
int32_t orc_halo_buffer(orc_ctx* h, void** ptr, uint64_t* record_bytes, uint64_t* n_records) {
  *ptr = h->c.halo.data();
  *record_bytes = sizeof(orc::AvoidRec);
  *n_records = h->c.halo.size();
  return LMB_OK;
}
int32_t orc_step_end(orc_ctx* h, const lmb_options* options, float dt, const lmb_agents_out* out) {
  if (!h || !options || !out) return LMB_ERR_INVALID_ARGUMENT;
  h->c.step_end(*options, dt, *out);
  return LMB_OK;
}
int32_t orc_pathing_results(orc_ctx* h, lmb_pathing_result* out, uint32_t capacity,
                            uint32_t* out_count) {
  *out_count = (uint32_t)h->c.pathing_results.size();
  if (*out_count > capacity) return LMB_ERR_CAPACITY;
  std::copy(h->c.pathing_results.begin(), h->c.pathing_results.end(), out);
  return LMB_OK;
}
int32_t orc_agent_path(orc_ctx* h, uint32_t slot, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t capacity, uint32_t* out_len) {
  *out_len = 0;
  auto it = h->c.agent_state.find(slot);
  if (it == h->c.agent_state.end() || !it->second.path || it->second.path_invalid) return LMB_OK;
  std::vector<uint32_t> nodes, portals;
  it->second.path->flatten(nodes, portals);
  *out_len = (uint32_t)nodes.size();
  if (nodes.size() > capacity) return LMB_ERR_CAPACITY;
  std::copy(nodes.begin(), nodes.end(), out_nodes);
  std::copy(portals.begin(), portals.end(), out_portals);
  return LMB_OK;
}

int32_t orc_agent_path_endpoints(orc_ctx* h, uint32_t slot, float* out_start_point, float* out_end_point) {
  auto it = h->c.agent_state.find(slot);
  if (it == h->c.agent_state.end() || !it->second.path) return LMB_ERR_INVALID_ARGUMENT;
  const Path& p = *it->second.path;
  out_start_point[0] = p.start_point.x; out_start_point[1] = p.start_point.y; out_start_point[2] = p.start_point.z;
  out_end_point[0] = p.end_point.x; out_end_point[1] = p.end_point.y; out_end_point[2] = p.end_point.z;
  return LMB_OK;
}
int32_t orc_agent_waypoints(orc_ctx* h, uint32_t capacity, uint32_t* out_kind, float* out_points,
                            uint32_t* out_link, uint32_t* out_count) {
  const auto& w = h->c.f_waypoints;
  *out_count = (uint32_t)w.size();
  if (w.size() > capacity) return LMB_ERR_CAPACITY;
  for (size_t i = 0; i < w.size(); i++) {
    out_kind[i] = w[i].kind;
    out_link[i] = w[i].link;
    float* p = out_points + 6 * i;
    p[0] = w[i].point.x; p[1] = w[i].point.y; p[2] = w[i].point.z;
    p[3] = w[i].end_point.x; p[4] = w[i].end_point.y; p[5] = w[i].end_point.z;
  }
  return LMB_OK;
}

int32_t orc_agent_avoidance_constraints(orc_ctx* h, uint32_t agent_index, uint32_t capacity, float* out_lines,
                                        uint32_t* out_n_original, uint32_t* out_n_fallback, uint32_t* out_ran) {
  *out_n_original = *out_n_fallback = *out_ran = 0;
  if (agent_index >= h->c.f_agents.size()) return LMB_ERR_INVALID_ARGUMENT;
  if (agent_index >= h->c.f_avoid_debug.size()) return LMB_OK;
  const AvoidDebugData& d = h->c.f_avoid_debug[agent_index];
  *out_ran = d.ran ? (d.fell_back ? 2 : 1) : 0;
  *out_n_original = (uint32_t)d.original.size();
  *out_n_fallback = (uint32_t)d.fallback.size();
  if (d.original.size() + d.fallback.size() > capacity) return LMB_ERR_CAPACITY;
  size_t k = 0;
  for (const auto& l : d.original) { for (int c = 0; c < 4; c++) out_lines[4 * k + c] = l[c]; k++; }
  for (const auto& l : d.fallback) { for (int c = 0; c < 4; c++) out_lines[4 * k + c] = l[c]; k++; }
  return LMB_OK;
}

int32_t orc_next_straight_point(orc_ctx* h, uint32_t path_len, const uint32_t* nodes,
                                const uint32_t* portals, uint32_t start_index,
                                const float* start_point, uint32_t end_index, const float* end_point,
                                uint32_t* out_index, uint32_t* out_kind, float* out_point) {
  Path path = Path::from_flat(h->c.nav, path_len, nodes, portals);
  auto r = path.find_next_point_in_straight_path(
      h->c.nav, path.from_flat_index(start_index), {start_point[0], start_point[1], start_point[2]},
      path.from_flat_index(end_index), {end_point[0], end_point[1], end_point[2]});
  *out_index = path.flat_index(r.first);
  *out_kind = r.second.is_link ? 1 : 0;
  out_point[0] = r.second.point.x;
  out_point[1] = r.second.point.y;
  out_point[2] = r.second.point.z;
  if (r.second.is_link) {
    out_point[3] = r.second.end_point.x;
    out_point[4] = r.second.end_point.y;
    out_point[5] = r.second.end_point.z;
  }
  return LMB_OK;
}

// Path::find_index_of_node / find_index_of_node_rev (path.rs:402-446) on a hand-made corridor, as a flat index.
int32_t orc_path_find_index_of_node(orc_ctx* h, uint32_t path_len, const uint32_t* nodes, const uint32_t* portals,
                                    uint32_t node, uint32_t reverse, uint32_t* out_found, uint32_t* out_index) {
  Path path = Path::from_flat(h->c.nav, path_len, nodes, portals);
  if (node >= h->c.nav.n_polys) return LMB_ERR_INVALID_ARGUMENT;
  auto r = reverse ? path.find_index_of_node_rev(h->c.nav, node) : path.find_index_of_node(h->c.nav, node);
  *out_found = r ? 1 : 0;
  *out_index = r ? path.flat_index(*r) : LMB_NONE;
  return LMB_OK;
}

// geometry.rs:176-189
int32_t orc_project_point_to_line_segment(const float* point, const float* start, const float* end,
                                          float* out_point, float* out_fraction) {
  auto r = project_point_to_line_segment({point[0], point[1], point[2]}, {start[0], start[1], start[2]},
                                         {end[0], end[1], end[2]});
  out_point[0] = r.first.x;
  out_point[1] = r.first.y;
  out_point[2] = r.first.z;
  *out_fraction = r.second;
  return LMB_OK;
}

int32_t orc_astar_adjacency(uint32_t n_states, const uint32_t* adj_begin, const float* adj_cost,
                            const int32_t* adj_action, const uint32_t* adj_state,
                            const float* heuristic, uint32_t start, uint32_t end,
                            int32_t* out_actions, uint32_t capacity, uint32_t* out_len,
                            uint32_t* out_found, uint32_t* out_explored) {
  AdjProblem p{n_states, adj_begin, adj_cost, adj_action, adj_state, heuristic, start, end};
  AStarResult r = astar_find_path(p);
  *out_found = r.found ? 1 : 0;
  *out_explored = r.explored_nodes;
  *out_len = (uint32_t)r.actions.size();
  if (r.actions.size() > capacity) return LMB_ERR_CAPACITY;
  for (size_t i = 0; i < r.actions.size(); i++) out_actions[i] = (int32_t)r.actions[i];
  return LMB_OK;
}

}  // extern "C"
