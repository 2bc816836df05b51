// TEST INFRASTRUCTURE — CPU oracle for landmass_b200 (see oracle/README.md).
//
// Phase D of `Archipelago::update`: local avoidance
// (crates/landmass/src/avoidance.rs:16-471).
//
// PARITY UNPINNED for the arithmetic in this file that belongs to the
// third-party crate `dodgy_2d = "0.5.4"` (crates/landmass/Cargo.toml:18):
// its source is NOT under /root/reference and cannot be fetched.  What is
// restated here is (1) landmass's own code (neighbour selection, border flood,
// loop stitching), which IS pinned by avoidance_test.rs:68-387, and (2) the
// published ORCA / RVO2 algorithm (van den Berg et al., "Reciprocal n-body
// collision avoidance"; RVO2 library linearProgram1/2/3) that dodgy_2d is a
// port of, with dodgy_2d's per-agent `avoidance_responsibility` as recalled in
// SURVEY.md Appendix C.  The reference's own velocity tests
// (avoidance_test.rs:389-853, tolerance 0.05) are the sanity floor.
//
// `VisibilitySet` (dodgy_2d::VisibilitySet) is restated as an exact angular
// occlusion structure: disjoint "cones" in a division-only pseudo-angle, each
// owned by the nearest line added so far.  A line is visible iff it owns a
// cone of non-zero width.  Canonical orders replace the reference's
// HashSet/HashMap iteration orders.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "landmass_oracle.h"
#include "orc_internal.h"

namespace orc {

// ---------------------------------------------------------------------------
// Visibility set
// ---------------------------------------------------------------------------
namespace {

// Monotone pseudo-angle in [0,4): 0 at +x, 1 at +y, 2 at -x, 3 at -y.
inline float pangle(V2 v) {
  float d = std::fabs(v.x) + std::fabs(v.y);
  float t = v.x / d;
  return v.y >= 0.0f ? 1.0f - t : 3.0f + t;
}
inline V2 pdir(float g) {
  if (g < 2.0f) {
    float t = 1.0f - g;
    return {t, 1.0f - std::fabs(t)};
  }
  float t = g - 3.0f;
  return {t, std::fabs(t) - 1.0f};
}
// distance along direction d (unnormalised) at which the ray from the origin
// meets the line through a,b
inline float ray_depth(V2 a, V2 b, V2 d) { return perp_dot(a, b) / perp_dot(d, b - a); }

struct Cone {
  float lo, hi;
  uint32_t line;
};

struct VisibilitySet {
  std::vector<V2> la, lb;
  std::vector<Cone> cones;  // sorted by lo, pairwise disjoint interiors

  // Tests (add_id < 0) or inserts (add_id = new line index) the span [lo,hi]
  // of line (a,b).  Returns whether any part is visible; when inserting,
  // `out` receives the updated cone list.
  bool process_piece(V2 a, V2 b, float lo, float hi, int add_id, std::vector<Cone>* out) const {
    bool visible = false;
    float cur = lo;
    bool tail_done = false;
    auto emit = [&](float l, float h, uint32_t line) {
      if (!out || !(l < h)) return;
      if (!out->empty() && out->back().line == line && out->back().hi == l)
        out->back().hi = h;
      else
        out->push_back({l, h, line});
    };
    for (const Cone& c : cones) {
      if (!(c.hi > lo && c.lo < hi)) {
        if (c.lo >= hi && !tail_done) {  // first cone past the span: the trailing gap goes before it
          tail_done = true;
          if (cur < hi) {
            visible = true;
            if (add_id >= 0) emit(cur, hi, (uint32_t)add_id);
          }
        }
        emit(c.lo, c.hi, c.line);
        continue;
      }
      if (c.lo > cur) {  // gap [cur, c.lo]
        visible = true;
        if (add_id >= 0) emit(cur, c.lo, (uint32_t)add_id);
        cur = c.lo;
      }
      float olo = std::fmax(cur, c.lo), ohi = std::fmin(hi, c.hi);
      if (c.lo < olo) emit(c.lo, olo, c.line);
      if (olo < ohi) {
        V2 d = pdir((olo + ohi) * 0.5f);
        float un = ray_depth(a, b, d), uc = ray_depth(la[c.line], lb[c.line], d);
        if (un < uc) {
          visible = true;
          emit(olo, ohi, add_id >= 0 ? (uint32_t)add_id : c.line);
        } else {
          emit(olo, ohi, c.line);
        }
      }
      if (ohi < c.hi) emit(ohi, c.hi, c.line);
      if (c.hi > cur) cur = c.hi;
    }
    if (!tail_done && cur < hi) {
      visible = true;
      if (add_id >= 0) emit(cur, hi, (uint32_t)add_id);
    }
    return visible;
  }

  bool process(V2 a, V2 b, bool add, uint32_t* out_id) {
    float det = perp_dot(a, b);
    if (det < 0.0f) std::swap(a, b);
    if (!(det != 0.0f)) {
      // The supporting line passes through the origin: zero angular width, so it can never
      // occlude (not added); a portal is still "visible" when the origin lies ON the segment
      // (agent standing on the portal, avoidance_test.rs:237-327).
      return !add && det == 0.0f && dot(a, b) <= 0.0f;
    }
    float lo = pangle(a), hi = pangle(b);
    if (lo == hi) return false;
    if (!add) {
      if (lo < hi) return process_piece(a, b, lo, hi, -1, nullptr);
      return process_piece(a, b, lo, 4.0f, -1, nullptr) || process_piece(a, b, 0.0f, hi, -1, nullptr);
    }
    int id = (int)la.size();
    la.push_back(a);  // registered so that pieces can be compared against it
    lb.push_back(b);
    bool visible;
    std::vector<Cone> nc;
    if (lo < hi) {
      visible = process_piece(a, b, lo, hi, id, &nc);
      if (visible) cones.swap(nc);
    } else {
      // span crosses the cut at pseudo-angle 0/4: two pieces, each committed if visible
      bool v1 = process_piece(a, b, lo, 4.0f, id, &nc);
      if (v1) cones.swap(nc);
      nc.clear();
      bool v2 = process_piece(a, b, 0.0f, hi, id, &nc);
      if (v2) cones.swap(nc);
      visible = v1 || v2;
    }
    if (!visible) {
      la.pop_back();
      lb.pop_back();
    } else if (out_id) {
      *out_id = (uint32_t)id;
    }
    return visible;
  }
  bool is_line_visible(V2 a, V2 b) { return process(a, b, false, nullptr); }
  bool add_line(V2 a, V2 b, uint32_t* id) { return process(a, b, true, id); }
  std::vector<uint32_t> get_visible_line_ids() const {  // ascending (canonical)
    std::vector<uint32_t> ids;
    for (const Cone& c : cones) ids.push_back(c.line);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    return ids;
  }
};

// std BinaryHeap<ExploreNode> with ExploreNode::partial_cmp = other.score.partial_cmp(self.score)
// (avoidance.rs:200-222):  a <= b  <=>  b.score <= a.score.
struct ExploreNode {
  uint32_t node;
  float score;
};
struct ExploreHeap {
  std::vector<ExploreNode> data;
  static bool le(const ExploreNode& a, const ExploreNode& b) { return b.score <= a.score; }
  void sift_up(size_t start, size_t pos) {
    ExploreNode hole = data[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (le(hole, data[parent])) break;
      data[pos] = data[parent];
      pos = parent;
    }
    data[pos] = hole;
  }
  void push(ExploreNode e) {
    data.push_back(e);
    sift_up(0, data.size() - 1);
  }
  ExploreNode pop() {
    ExploreNode item = data.back();
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      size_t end = data.size(), pos = 0, child = 1;
      ExploreNode hole = data[0];
      while (end >= 2 && child <= end - 2) {
        if (le(data[child], data[child + 1])) child += 1;
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
    return item;
  }
};

}  // namespace

// nav_mesh_borders_to_dodgy_obstacles (avoidance.rs:189-471)
std::vector<Obstacle> nav_mesh_borders_to_dodgy_obstacles(const NavData& nav, V3 agent_point3,
                                                          uint32_t agent_node, float distance_limit) {
  distance_limit = distance_limit * distance_limit;
  VisibilitySet vis;
  // vertex identity: (kind 0 = mesh vertex (global index), kind 1 = new vertex)
  typedef std::pair<uint32_t, uint32_t> VId;
  std::unordered_map<uint32_t, std::pair<VId, VId>> border_edges;
  std::vector<V2> new_vertices;
  std::unordered_set<uint32_t> explored;
  ExploreHeap next_nodes;
  next_nodes.push({agent_node, 0.0f});
  V2 agent_point = xy(agent_point3);

  auto vertex_rel = [&](uint32_t node, uint32_t gv, V2 rel) { return xy(nav.world_vertex(node, gv)) - rel; };

  while (!next_nodes.data.empty()) {
    uint32_t node = next_nodes.pop().node;
    if (!explored.insert(node).second) continue;
    uint32_t eb = nav.poly_edge_begin[node], n = nav.poly_edge_begin[node + 1] - eb;
    auto consider = [&](V2 v1, V2 v2, uint32_t dst) {
      if (!vis.is_line_visible(v1, v2)) return;
      float score = std::fmin(length_squared(v1), length_squared(v2));
      if (score < distance_limit) next_nodes.push({dst, score});
    };
    for (uint32_t e = 0; e < n; e++) {
      uint32_t nb = nav.edge_neighbor[eb + e];
      if (nb == LMB_NONE) continue;
      uint32_t l, r;
      nav.edge_indices(node, e, l, r);  // (vertex_1, vertex_2) = (left, right)
      consider(vertex_rel(node, r, agent_point), vertex_rel(node, l, agent_point), nb);
    }
    for (uint32_t k = nav.node_link_begin[node]; k < nav.node_link_begin[node + 1]; k++) {
      uint32_t lk = nav.node_links[k];
      if (nav.link_kind[lk] != LMB_LINK_BOUNDARY) continue;
      consider(xy(nav.link_portal[lk].second) - agent_point, xy(nav.link_portal[lk].first) - agent_point,
               nav.link_dst_node[lk]);
    }
    auto mit = nav.modified.find(node);
    if (mit != nav.modified.end()) {
      const Island& is = nav.islands[nav.poly_island[node]];
      uint32_t island_verts = is.vert_end - is.vert_begin;
      auto resolve = [&](uint32_t index, V2& pt, VId& id) {
        if (index >= island_verts) {
          V2 nv = mit->second.new_vertices[index - island_verts];
          id = {1u, (uint32_t)new_vertices.size()};
          new_vertices.push_back(nv);
          pt = nv - agent_point;
        } else {
          pt = vertex_rel(node, is.vert_begin + index, agent_point);
          id = {0u, is.vert_begin + index};
        }
      };
      for (auto& lr : mit->second.new_boundary) {
        V2 lp, rp;
        VId li, ri;
        resolve(lr.first, lp, li);
        resolve(lr.second, rp, ri);
        uint32_t id;
        if (vis.add_line(lp, rp, &id)) border_edges[id] = {li, ri};
      }
    } else {
      for (uint32_t e = 0; e < n; e++) {  // remaining (border) edges, ascending
        if (nav.edge_neighbor[eb + e] != LMB_NONE) continue;
        uint32_t bv2, bv1;
        nav.edge_indices(node, e, bv2, bv1);
        uint32_t id;
        if (vis.add_line(vertex_rel(node, bv1, agent_point), vertex_rel(node, bv2, agent_point), &id))
          border_edges[id] = {VId{0u, bv1}, VId{0u, bv2}};
      }
    }
  }

  // stitch visible edges into loops (avoidance.rs:389-439)
  std::vector<std::vector<VId>> finished, unfinished;
  for (uint32_t line_id : vis.get_visible_line_ids()) {
    auto edge = border_edges.at(line_id);
    int left_loop = -1, right_loop = -1;
    for (size_t li = 0; li < unfinished.size(); li++) {
      auto& loop = unfinished[li];
      if (left_loop < 0 && loop.back() == edge.first) {
        left_loop = (int)li;
        if (right_loop >= 0) break;
      }
      if (right_loop < 0 && loop.front() == edge.second) {
        right_loop = (int)li;
        if (left_loop >= 0) break;
      }
    }
    if (left_loop < 0 && right_loop < 0) {
      unfinished.push_back({edge.first, edge.second});
    } else if (left_loop >= 0 && right_loop < 0) {
      unfinished[left_loop].push_back(edge.second);
    } else if (left_loop < 0) {
      unfinished[right_loop].insert(unfinished[right_loop].begin(), edge.first);
    } else if (left_loop == right_loop) {
      finished.push_back(unfinished[left_loop]);
      unfinished.erase(unfinished.begin() + left_loop);
    } else {
      // Vec::swap_remove twice, then push(left ++ right)
      auto swap_remove = [&](size_t i) {
        std::vector<VId> v = std::move(unfinished[i]);
        unfinished[i] = std::move(unfinished.back());
        unfinished.pop_back();
        return v;
      };
      std::vector<VId> left = swap_remove((size_t)left_loop);
      size_t ri = ((size_t)right_loop == unfinished.size()) ? (size_t)left_loop : (size_t)right_loop;
      std::vector<VId> right = swap_remove(ri);
      left.insert(left.end(), right.begin(), right.end());
      unfinished.push_back(std::move(left));
    }
  }
  auto to_point = [&](VId id) -> V2 {
    if (id.first == 1u) return new_vertices[id.second];
    // vertex_index_to_dodgy_vec(island, index, Vec2::ZERO): world xy minus zero.
    // The node is only needed to find the island; find it from the vertex range.
    uint32_t island = 0;
    while (!(id.second >= nav.islands[island].vert_begin && id.second < nav.islands[island].vert_end)) island++;
    return xy(nav.islands[island].xf.apply(nav.vertices[id.second])) - V2{0.0f, 0.0f};
  };
  std::vector<Obstacle> result;
  for (auto& loop : finished) {
    Obstacle o;
    o.closed = true;
    for (size_t i = loop.size(); i-- > 0;) o.vertices.push_back(to_point(loop[i]));
    result.push_back(std::move(o));
  }
  for (auto& loop : unfinished) {
    Obstacle o;
    o.closed = false;
    for (size_t i = loop.size(); i-- > 0;) o.vertices.push_back(to_point(loop[i]));
    result.push_back(std::move(o));
  }
  return result;
}

// ---------------------------------------------------------------------------
// ORCA (dodgy_2d restatement from RVO2; see header comment)
// ---------------------------------------------------------------------------
namespace {
constexpr float RVO_EPSILON = 0.00001f;

struct Line {
  V2 point, direction;
};
struct DodgyAgent {
  V2 position, velocity;
  float radius, responsibility;
};

inline float det2(V2 a, V2 b) { return perp_dot(a, b); }

// one neighbour -> one half-plane
Line line_for_neighbour(const DodgyAgent& self, const DodgyAgent& other, float time_horizon,
                        float time_step) {
  V2 relative_position = other.position - self.position;
  V2 relative_velocity = self.velocity - other.velocity;
  float dist_sq = length_squared(relative_position);
  float combined_radius = self.radius + other.radius;
  float combined_radius_sq = combined_radius * combined_radius;
  Line line;
  V2 u;
  bool inside_vo;
  if (dist_sq > combined_radius_sq) {
    float inv_time_horizon = 1.0f / time_horizon;
    V2 w = relative_velocity - relative_position * inv_time_horizon;
    float w_length_sq = length_squared(w);
    float dot1 = dot(w, relative_position);
    if (dot1 < 0.0f && dot1 * dot1 > combined_radius_sq * w_length_sq) {
      // project on the cut-off circle
      float w_length = std::sqrt(w_length_sq);
      V2 unit_w = w / w_length;
      line.direction = {unit_w.y, -unit_w.x};
      float cutoff_radius = combined_radius * inv_time_horizon;
      u = unit_w * (cutoff_radius - w_length);
      inside_vo = w_length_sq < cutoff_radius * cutoff_radius;
    } else {
      // project on the legs
      float leg = std::sqrt(dist_sq - combined_radius_sq);
      // dodgy_2d picks the leg by `determinant(..).signum()`: +0.0 counts as positive, so an exactly
      // symmetric approach takes the LEFT leg (textbook RVO2 tests `> 0` and goes right) — pinned by
      // lib_test.rs:426-486 (`agent_speeds_up_to_avoid_character`)
      if (!std::signbit(det2(relative_position, w))) {
        line.direction = V2{relative_position.x * leg - relative_position.y * combined_radius,
                            relative_position.x * combined_radius + relative_position.y * leg} /
                         dist_sq;
      } else {
        line.direction = -(V2{relative_position.x * leg + relative_position.y * combined_radius,
                              -relative_position.x * combined_radius + relative_position.y * leg} /
                           dist_sq);
      }
      float dot2 = dot(relative_velocity, line.direction);
      u = line.direction * dot2 - relative_velocity;
      inside_vo = det2(line.direction, relative_velocity) < 0.0f;
    }
  } else {
    // already colliding: cut-off circle of the time step
    float inv_time_step = 1.0f / time_step;
    V2 w = relative_velocity - relative_position * inv_time_step;
    float w_length = length(w);
    V2 unit_w = w / w_length;
    line.direction = {unit_w.y, -unit_w.x};
    float cutoff_radius = combined_radius * inv_time_step;
    u = unit_w * (cutoff_radius - w_length);
    inside_vo = w_length < cutoff_radius;
  }
  float responsibility =
      inside_vo ? self.responsibility / (self.responsibility + other.responsibility) : 1.0f;
  line.point = self.velocity + u * responsibility;
  return line;
}

struct ObVertex {
  V2 point;
  bool convex;
  bool has_prev, has_next;
  V2 unit_dir;       // towards next (valid if has_next)
  V2 prev_unit_dir;  // prev -> this (valid if has_prev)
};

void lines_for_obstacle(const DodgyAgent& self, const Obstacle& ob, float margin, float time_horizon,
                        std::vector<Line>& lines) {
  size_t n = ob.vertices.size();
  if (n < 2) return;
  // Clearance from obstacle edges is the obstacle margin alone, NOT the agent radius: landmass
  // passes obstacle_margin = 0 because "the nav mesh is the valid region" (avoidance.rs:143-146)
  // — an agent centre may touch a border.  Pinned by lib_test.rs:156-423, where agents of radius
  // 0.5 in a 1-wide corridor keep their full desired velocity (tolerance 1e-7).
  (void)self.radius;
  float radius = margin;
  float inv_th = 1.0f / time_horizon;
  std::vector<ObVertex> vs(n);
  for (size_t i = 0; i < n; i++) {
    ObVertex& v = vs[i];
    v.point = ob.vertices[i];
    v.has_prev = ob.closed || i > 0;
    v.has_next = ob.closed || i + 1 < n;
    V2 prev = ob.vertices[(i + n - 1) % n], next = ob.vertices[(i + 1) % n];
    if (v.has_next) v.unit_dir = normalize(next - v.point);
    if (v.has_prev) v.prev_unit_dir = normalize(v.point - prev);
    // left_of(prev, this, next) >= 0; chain ends count as convex
    v.convex = (v.has_prev && v.has_next && n > 2) ? det2(prev - next, v.point - prev) >= 0.0f : true;
  }
  size_t n_edges = ob.closed ? n : n - 1;
  for (size_t e = 0; e < n_edges; e++) {
    const ObVertex* o1 = &vs[e];
    const ObVertex* o2 = &vs[(e + 1) % n];
    V2 rel1 = o1->point - self.position, rel2 = o2->point - self.position;
    // only edges whose solid side faces away: the agent must be on the right of o1->o2
    if (!(det2(o2->point - o1->point, self.position - o1->point) < 0.0f)) continue;
    bool covered = false;
    for (const Line& l : lines) {
      if (det2(rel1 * inv_th - l.point, l.direction) - inv_th * radius >= -RVO_EPSILON &&
          det2(rel2 * inv_th - l.point, l.direction) - inv_th * radius >= -RVO_EPSILON) {
        covered = true;
        break;
      }
    }
    if (covered) continue;
    float dist_sq1 = length_squared(rel1), dist_sq2 = length_squared(rel2);
    float radius_sq = radius * radius;
    V2 obstacle_vector = o2->point - o1->point;
    float s = dot(-rel1, obstacle_vector) / length_squared(obstacle_vector);
    float dist_sq_line = length_squared(-rel1 - obstacle_vector * s);
    Line line;
    if (s < 0.0f && dist_sq1 <= radius_sq) {
      if (o1->convex) {
        line.point = {0.0f, 0.0f};
        line.direction = normalize(V2{-rel1.y, rel1.x});
        lines.push_back(line);
      }
      continue;
    } else if (s > 1.0f && dist_sq2 <= radius_sq) {
      if (o2->convex && (!o2->has_next || det2(rel2, o2->unit_dir) >= 0.0f)) {
        line.point = {0.0f, 0.0f};
        line.direction = normalize(V2{-rel2.y, rel2.x});
        lines.push_back(line);
      }
      continue;
    } else if (s >= 0.0f && s <= 1.0f && dist_sq_line <= radius_sq) {
      line.point = {0.0f, 0.0f};
      line.direction = -o1->unit_dir;
      lines.push_back(line);
      continue;
    }
    V2 left_leg, right_leg;
    V2 edge_dir = o1->unit_dir;
    if (s < 0.0f && dist_sq_line <= radius_sq) {
      if (!o1->convex) continue;
      o2 = o1;
      float leg1 = std::sqrt(dist_sq1 - radius_sq);
      left_leg = V2{rel1.x * leg1 - rel1.y * radius, rel1.x * radius + rel1.y * leg1} / dist_sq1;
      right_leg = V2{rel1.x * leg1 + rel1.y * radius, -rel1.x * radius + rel1.y * leg1} / dist_sq1;
    } else if (s > 1.0f && dist_sq_line <= radius_sq) {
      if (!o2->convex) continue;
      o1 = o2;
      float leg2 = std::sqrt(dist_sq2 - radius_sq);
      left_leg = V2{rel2.x * leg2 - rel2.y * radius, rel2.x * radius + rel2.y * leg2} / dist_sq2;
      right_leg = V2{rel2.x * leg2 + rel2.y * radius, -rel2.x * radius + rel2.y * leg2} / dist_sq2;
    } else {
      if (o1->convex) {
        float leg1 = std::sqrt(dist_sq1 - radius_sq);
        left_leg = V2{rel1.x * leg1 - rel1.y * radius, rel1.x * radius + rel1.y * leg1} / dist_sq1;
      } else {
        left_leg = -edge_dir;
      }
      if (o2->convex) {
        float leg2 = std::sqrt(dist_sq2 - radius_sq);
        right_leg = V2{rel2.x * leg2 + rel2.y * radius, -rel2.x * radius + rel2.y * leg2} / dist_sq2;
      } else {
        right_leg = edge_dir;
      }
    }
    rel1 = o1->point - self.position;
    rel2 = o2->point - self.position;
    bool left_foreign = false, right_foreign = false;
    if (o1->convex && o1->has_prev && det2(left_leg, -o1->prev_unit_dir) >= 0.0f) {
      left_leg = -o1->prev_unit_dir;
      left_foreign = true;
    }
    if (o2->convex && o2->has_next && det2(right_leg, o2->unit_dir) <= 0.0f) {
      right_leg = o2->unit_dir;
      right_foreign = true;
    }
    V2 left_cutoff = rel1 * inv_th, right_cutoff = rel2 * inv_th;
    V2 cutoff_vec = right_cutoff - left_cutoff;
    bool same = (o1 == o2);
    float t = same ? 0.5f : dot(self.velocity - left_cutoff, cutoff_vec) / length_squared(cutoff_vec);
    float t_left = dot(self.velocity - left_cutoff, left_leg);
    float t_right = dot(self.velocity - right_cutoff, right_leg);
    if ((t < 0.0f && t_left < 0.0f) || (same && t_left < 0.0f && t_right < 0.0f)) {
      V2 unit_w = normalize(self.velocity - left_cutoff);
      line.direction = {unit_w.y, -unit_w.x};
      line.point = left_cutoff + unit_w * (radius * inv_th);
      lines.push_back(line);
      continue;
    } else if (t > 1.0f && t_right < 0.0f) {
      V2 unit_w = normalize(self.velocity - right_cutoff);
      line.direction = {unit_w.y, -unit_w.x};
      line.point = right_cutoff + unit_w * (radius * inv_th);
      lines.push_back(line);
      continue;
    }
    const float INF = std::numeric_limits<float>::infinity();
    float dist_sq_cutoff = (t < 0.0f || t > 1.0f || same)
                               ? INF
                               : length_squared(self.velocity - (left_cutoff + cutoff_vec * t));
    float dist_sq_left =
        t_left < 0.0f ? INF : length_squared(self.velocity - (left_cutoff + left_leg * t_left));
    float dist_sq_right =
        t_right < 0.0f ? INF : length_squared(self.velocity - (right_cutoff + right_leg * t_right));
    if (dist_sq_cutoff <= dist_sq_left && dist_sq_cutoff <= dist_sq_right) {
      line.direction = -edge_dir;
      line.point = left_cutoff + V2{-line.direction.y, line.direction.x} * (radius * inv_th);
      lines.push_back(line);
    } else if (dist_sq_left <= dist_sq_right) {
      if (left_foreign) continue;
      line.direction = left_leg;
      line.point = left_cutoff + V2{-line.direction.y, line.direction.x} * (radius * inv_th);
      lines.push_back(line);
    } else {
      if (right_foreign) continue;
      line.direction = -right_leg;
      line.point = right_cutoff + V2{-line.direction.y, line.direction.x} * (radius * inv_th);
      lines.push_back(line);
    }
  }
}

bool linear_program_1(const std::vector<Line>& lines, size_t line_no, float radius, V2 opt_velocity,
                      bool direction_opt, V2& result) {
  const Line& ln = lines[line_no];
  float dot_product = dot(ln.point, ln.direction);
  float discriminant = dot_product * dot_product + radius * radius - length_squared(ln.point);
  if (discriminant < 0.0f) return false;
  float sqrt_discriminant = std::sqrt(discriminant);
  float t_left = -dot_product - sqrt_discriminant;
  float t_right = -dot_product + sqrt_discriminant;
  for (size_t i = 0; i < line_no; i++) {
    float denominator = det2(ln.direction, lines[i].direction);
    float numerator = det2(lines[i].direction, ln.point - lines[i].point);
    if (std::fabs(denominator) <= RVO_EPSILON) {
      if (numerator < 0.0f) return false;
      continue;
    }
    float t = numerator / denominator;
    if (denominator >= 0.0f)
      t_right = std::fmin(t_right, t);
    else
      t_left = std::fmax(t_left, t);
    if (t_left > t_right) return false;
  }
  if (direction_opt) {
    if (dot(opt_velocity, ln.direction) > 0.0f)
      result = ln.point + ln.direction * t_right;
    else
      result = ln.point + ln.direction * t_left;
  } else {
    float t = dot(ln.direction, opt_velocity - ln.point);
    if (t < t_left)
      result = ln.point + ln.direction * t_left;
    else if (t > t_right)
      result = ln.point + ln.direction * t_right;
    else
      result = ln.point + ln.direction * t;
  }
  return true;
}

size_t linear_program_2(const std::vector<Line>& lines, float radius, V2 opt_velocity,
                        bool direction_opt, V2& result) {
  if (direction_opt) {
    result = opt_velocity * radius;
  } else if (length_squared(opt_velocity) > radius * radius) {
    result = normalize(opt_velocity) * radius;
  } else {
    result = opt_velocity;
  }
  for (size_t i = 0; i < lines.size(); i++) {
    if (det2(lines[i].direction, lines[i].point - result) > 0.0f) {
      V2 temp = result;
      if (!linear_program_1(lines, i, radius, opt_velocity, direction_opt, result)) {
        result = temp;
        return i;
      }
    }
  }
  return lines.size();
}

// `last_proj` (optional) receives the projected constraint set of the last re-solve: the
// Fallback.fallback_constraints of the debug-avoidance export (debug.rs:470-481).
void linear_program_3(const std::vector<Line>& lines, size_t num_obst_lines, size_t begin_line,
                      float radius, V2& result, std::vector<Line>* last_proj = nullptr) {
  float distance_ = 0.0f;
  for (size_t i = begin_line; i < lines.size(); i++) {
    if (det2(lines[i].direction, lines[i].point - result) > distance_) {
      std::vector<Line> proj(lines.begin(), lines.begin() + num_obst_lines);
      for (size_t j = num_obst_lines; j < i; j++) {
        Line line;
        float determinant = det2(lines[i].direction, lines[j].direction);
        if (std::fabs(determinant) <= RVO_EPSILON) {
          if (dot(lines[i].direction, lines[j].direction) > 0.0f) continue;
          line.point = (lines[i].point + lines[j].point) * 0.5f;
        } else {
          line.point = lines[i].point +
                       lines[i].direction *
                           (det2(lines[j].direction, lines[i].point - lines[j].point) / determinant);
        }
        line.direction = normalize(lines[j].direction - lines[i].direction);
        proj.push_back(line);
      }
      V2 temp = result;
      if (linear_program_2(proj, radius, V2{-lines[i].direction.y, lines[i].direction.x}, true,
                           result) < proj.size())
        result = temp;
      distance_ = det2(lines[i].direction, lines[i].point - result);
      if (last_proj) *last_proj = proj;
    }
  }
}

V2 compute_avoiding_velocity(const DodgyAgent& self, const std::vector<DodgyAgent>& neighbours,
                             const std::vector<Obstacle>& obstacles, V2 preferred_velocity,
                             float max_speed, float time_step, float obstacle_margin,
                             float time_horizon, float obstacle_time_horizon,
                             AvoidDebugData* debug = nullptr) {
  std::vector<Line> lines;
  for (const Obstacle& ob : obstacles)
    lines_for_obstacle(self, ob, obstacle_margin, obstacle_time_horizon, lines);
  size_t obstacle_line_count = lines.size();
  for (const DodgyAgent& nb : neighbours)
    lines.push_back(line_for_neighbour(self, nb, time_horizon, time_step));
  V2 result;
  size_t fail = linear_program_2(lines, max_speed, preferred_velocity, false, result);
  std::vector<Line> last_proj;
  if (fail < lines.size())
    linear_program_3(lines, obstacle_line_count, fail, max_speed, result, debug ? &last_proj : nullptr);
  if (debug) {
    // compute_avoiding_velocity_with_debug (avoidance.rs:163-171)
    debug->ran = true;
    debug->fell_back = fail < lines.size();
    debug->original.clear();
    debug->fallback.clear();
    for (const Line& l : lines) debug->original.push_back({l.point.x, l.point.y, l.direction.x, l.direction.y});
    for (const Line& l : last_proj) debug->fallback.push_back({l.point.x, l.point.y, l.direction.x, l.direction.y});
  }
  return result;
}

// 3-D kd-tree with radius query, results ascending by (distance, index) —
// kdtree 0.7 `within` behaviour.
struct KdTree3 {
  struct Pt {
    float p[3];
    uint32_t id;
  };
  std::vector<Pt> pts;
  void build(size_t lo, size_t hi, int axis) {
    if (hi - lo <= 8) return;
    size_t mid = (lo + hi) / 2;
    std::nth_element(pts.begin() + lo, pts.begin() + mid, pts.begin() + hi,
                     [axis](const Pt& a, const Pt& b) { return a.p[axis] < b.p[axis]; });
    build(lo, mid, (axis + 1) % 3);
    build(mid + 1, hi, (axis + 1) % 3);
  }
  void finish() { build(0, pts.size(), 0); }
  void query(size_t lo, size_t hi, int axis, const float* q, float r2, float r,
             std::vector<std::pair<float, uint32_t>>& out) const {
    if (hi - lo <= 8) {
      for (size_t i = lo; i < hi; i++) {
        float dx = pts[i].p[0] - q[0], dy = pts[i].p[1] - q[1], dz = pts[i].p[2] - q[2];
        // squared_euclidean: sum of squared differences in dimension order
        float d = dx * dx + dy * dy + dz * dz;
        if (d <= r2) out.push_back({d, pts[i].id});
      }
      return;
    }
    size_t mid = (lo + hi) / 2;
    float delta = q[axis] - pts[mid].p[axis];
    {
      float dx = pts[mid].p[0] - q[0], dy = pts[mid].p[1] - q[1], dz = pts[mid].p[2] - q[2];
      float d = dx * dx + dy * dy + dz * dz;
      if (d <= r2) out.push_back({d, pts[mid].id});
    }
    // r (not r2) with slack so that rounding never prunes a qualifying point
    float slack = r * 1.0001f + 1e-6f;
    if (delta <= slack) query(lo, mid, (axis + 1) % 3, q, r2, r, out);
    if (delta >= -slack) query(mid + 1, hi, (axis + 1) % 3, q, r2, r, out);
  }
  std::vector<std::pair<float, uint32_t>> within(const float* q, float r2) const {
    std::vector<std::pair<float, uint32_t>> out;
    query(0, pts.size(), 0, q, r2, std::sqrt(r2), out);
    std::sort(out.begin(), out.end());
    return out;
  }
};

}  // namespace

// Per-agent avoidance record: exactly what one agent contributes to every other agent's
// avoidance (avoidance.rs:41-57).  Same 32-byte layout as the device halo record.
void build_avoid_records(const std::vector<AgentIn>& agents, const std::vector<AgentPersist*>& persist,
                         const std::vector<Sampled>& agent_node, const lmb_options& opt, AvoidRec* out) {
  for (size_t i = 0; i < agents.size(); i++) {
    AvoidRec r{};
    r.valid = agent_node[i].valid ? 1u : 0u;
    r.p = agent_node[i].point;
    r.v = xy(agents[i].velocity);
    r.radius = agents[i].radius;
    r.responsibility = persist[i]->state == LMB_STATE_REACHED_TARGET
                           ? opt.reached_destination_avoidance_responsibility
                           : 1.0f;
    out[i] = r;
  }
}

// apply_avoidance_to_agents (avoidance.rs:16-179).  `records` holds ALL agents of the
// archipelago (all ranks); this call solves for the agents [first, first + agents.size()).
void apply_avoidance_to_agents(const NavData& nav, const std::vector<AvoidRec>& records, uint32_t first,
                               const std::vector<AgentIn>& agents,
                               const std::vector<AgentPersist*>& persist,
                               const std::vector<Sampled>& agent_node,
                               const std::vector<CharacterIn>& characters, const lmb_options& opt,
                               float delta_time, std::vector<AvoidDebugData>* debug) {
  if (delta_time == 0.0f) delta_time = 1.0f;
  const size_t n = agents.size();
  KdTree3 agent_tree, character_tree;
  float agent_max_radius = 0.0f;
  for (size_t g = 0; g < records.size(); g++) {
    const AvoidRec& r = records[g];
    if (!r.valid) continue;
    agent_tree.pts.push_back({{r.p.x, r.p.y, r.p.z}, (uint32_t)g});
    agent_max_radius = std::fmax(agent_max_radius, r.radius);
  }
  agent_tree.finish();
  std::vector<DodgyAgent> dodgy_chars(characters.size());
  for (size_t i = 0; i < characters.size(); i++) {
    if (!characters[i].valid) continue;
    V3 p = characters[i].point;
    dodgy_chars[i] = {xy(p), xy(characters[i].velocity), characters[i].radius, 0.0f};
    character_tree.pts.push_back({{p.x, p.y, p.z}, (uint32_t)i});
  }
  character_tree.finish();
  auto dodgy_of = [&](uint32_t g) {
    const AvoidRec& r = records[g];
    return DodgyAgent{xy(r.p), r.v, r.radius, r.responsibility};
  };

  float neighbourhood = agent_max_radius + opt.neighbourhood;
  float neighbourhood_squared = neighbourhood * neighbourhood;
  for (size_t i = 0; i < n; i++) {
    if (!agent_node[i].valid) continue;
    const uint32_t self_g = first + (uint32_t)i;
    V3 ap = agent_node[i].point;
    float q[3] = {ap.x, ap.y, ap.z};
    auto nearby_agents = agent_tree.within(q, neighbourhood_squared);
    auto nearby_chars = character_tree.within(q, neighbourhood_squared);
    std::vector<DodgyAgent> neighbours;
    for (auto& dn : nearby_agents) {
      if (dn.second == self_g) continue;
      DodgyAgent d = dodgy_of(dn.second);
      float nh = opt.neighbourhood + d.radius;
      if (dn.first < nh * nh) neighbours.push_back(d);
    }
    for (auto& dn : nearby_chars)
      if (dn.first < neighbourhood_squared) neighbours.push_back(dodgy_chars[dn.second]);
    std::vector<Obstacle> obstacles =
        nav_mesh_borders_to_dodgy_obstacles(nav, ap, agent_node[i].node, opt.neighbourhood);
    V2 preferred = xy(persist[i]->desired_move);
    V2 v = compute_avoiding_velocity(dodgy_of(self_g), neighbours, obstacles, preferred, agents[i].max_speed,
                                     delta_time, 0.0f, opt.avoidance_time_horizon,
                                     opt.obstacle_avoidance_time_horizon,
                                     debug && (agents[i].flags & LMB_AGENT_KEEP_AVOIDANCE_DATA) ? &(*debug)[i] : nullptr);
    persist[i]->desired_move = {v.x, v.y, 0.0f};
  }
}

}  // namespace orc

using namespace orc;

extern "C" {

int32_t orc_borders_to_obstacles(orc_ctx* h, const float* point, uint32_t node, float distance_limit,
                                 uint32_t* out_n_obstacles, uint32_t* out_vert_begin,
                                 uint32_t* out_closed, float* out_verts, uint32_t obstacle_capacity,
                                 uint32_t vert_capacity) {
  auto obs = nav_mesh_borders_to_dodgy_obstacles(h->c.nav, {point[0], point[1], point[2]}, node,
                                                 distance_limit);
  *out_n_obstacles = (uint32_t)obs.size();
  if (obs.size() > obstacle_capacity) return LMB_ERR_CAPACITY;
  uint32_t cursor = 0;
  for (size_t i = 0; i < obs.size(); i++) {
    out_vert_begin[i] = cursor;
    out_closed[i] = obs[i].closed ? 1 : 0;
    if (cursor + obs[i].vertices.size() > vert_capacity) return LMB_ERR_CAPACITY;
    for (auto& v : obs[i].vertices) {
      out_verts[2 * cursor] = v.x;
      out_verts[2 * cursor + 1] = v.y;
      cursor++;
    }
  }
  out_vert_begin[obs.size()] = cursor;
  return LMB_OK;
}

int32_t orc_compute_avoiding_velocity(const float* agent, uint32_t n_neighbours,
                                      const float* neighbours, uint32_t n_obstacles,
                                      const uint32_t* obstacle_vert_begin,
                                      const uint32_t* obstacle_closed, const float* obstacle_verts,
                                      const float* preferred_velocity, float max_speed,
                                      float time_step, float obstacle_margin, float time_horizon,
                                      float obstacle_time_horizon, float* out_velocity) {
  auto mk = [](const float* a) { return DodgyAgent{{a[0], a[1]}, {a[2], a[3]}, a[4], a[5]}; };
  DodgyAgent self = mk(agent);
  std::vector<DodgyAgent> nbs;
  for (uint32_t i = 0; i < n_neighbours; i++) nbs.push_back(mk(neighbours + 6 * i));
  std::vector<Obstacle> obs(n_obstacles);
  for (uint32_t i = 0; i < n_obstacles; i++) {
    obs[i].closed = obstacle_closed[i] != 0;
    for (uint32_t v = obstacle_vert_begin[i]; v < obstacle_vert_begin[i + 1]; v++)
      obs[i].vertices.push_back({obstacle_verts[2 * v], obstacle_verts[2 * v + 1]});
  }
  V2 r = compute_avoiding_velocity(self, nbs, obs, {preferred_velocity[0], preferred_velocity[1]},
                                   max_speed, time_step, obstacle_margin, time_horizon,
                                   obstacle_time_horizon);
  out_velocity[0] = r.x;
  out_velocity[1] = r.y;
  return LMB_OK;
}
}
