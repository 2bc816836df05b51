// landmass_b200_linking.hpp — C++17 mirror of the island-linking part of `NavigationData::update`
// (nav_data.rs:326-514 boundary links, 827-1011 modified nodes, 1089-1156 region connectivity) and of the
// flattening of a world of several islands into `lmb_navdata_desc`.  Companion of landmass_b200.hpp; the same
// restatement as landmass_b200/linking.py + nav_data.flatten (Python), checked against it array by array in
// tests/test_host_mirror_cpp.py.  Not mirrored here: animation links.
//
// Canonical orders replace the reference's HashMap / HashSet iteration orders exactly as in the Python mirror
// (SURVEY.md Appendix E): islands in the order given (= DenseSlotMap order), boundary edges by (polygon, edge),
// links sorted by (source island, source polygon, creation order) with a link's reverse right behind it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "landmass_b200.hpp"

namespace landmass_b200 {

struct Island {  // island.rs:16-27
  Transform transform;
  ValidNavigationMesh mesh;
};

class NavigationData {
 public:
  // edge_link_distance: lib.rs:273-276 (0.01 in `Archipelago::update`)
  explicit NavigationData(std::vector<Island> islands, std::vector<std::pair<uint32_t, float>> type_index_costs = {},
                          float edge_link_distance = 0.01f)
      : islands_(std::move(islands)) {
    std::sort(type_index_costs.begin(), type_index_costs.end(),
              [](const auto& a, const auto& b) { return a.first < b.first; });
    for (const auto& kv : type_index_costs) cost_index_.push_back(kv.first), cost_value_.push_back(kv.second);
    build_world();
    link_islands(edge_link_distance);
    modified_nodes(edge_link_distance);
    update_regions();
    flatten();
  }
  NavigationData(const NavigationData&) = delete;
  NavigationData& operator=(const NavigationData&) = delete;
  const lmb_navdata_desc& desc() const { return desc_; }
  uint32_t node_id(uint32_t island, uint32_t polygon) const { return poly_begin_[island] + polygon; }
  size_t n_links() const { return links_.size(); }
  // what identifies a link across rebuilds (the reference keeps the OffMeshLinkId of a link whose islands did not
  // change, nav_data.rs:352-403): its end nodes as (island index, polygon) and its portal
  struct LinkInfo {
    uint32_t src_island, src_polygon, dst_island, dst_polygon;
    std::array<float, 6> portal;
  };
  LinkInfo link_info(size_t l) const {
    const Link& k = links_[l];
    return {k.src_island, k.src_polygon, k.dst_island, k.dst_polygon, {k.p0.x, k.p0.y, k.p0.z, k.p1.x, k.p1.y, k.p1.z}};
  }

 private:
  struct V3d {
    float x, y, z;
  };
  struct Link {  // OffMeshLink::BoundaryLink (nav_data.rs:78-117)
    uint32_t src_island, src_polygon, dst_island, dst_polygon, dst_type;
    V3d p0, p1;
    size_t pair;  // creation index of the reverse link, later its final index
  };
  struct Modified {  // nav_data.rs:119-132
    uint32_t island, polygon;
    std::vector<std::pair<uint32_t, uint32_t>> new_boundary;
    std::vector<std::pair<float, float>> new_vertices;
  };

  static float dot3(V3d a, V3d b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
  static V3d sub(V3d a, V3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
  static V3d add(V3d a, V3d b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
  static V3d mul(V3d a, float s) { return {a.x * s, a.y * s, a.z * s}; }
  static float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }  // NaN passes through

  // transform.apply(vertex) for every vertex (util.rs:358-361; glam's quaternion product order)
  void build_world() {
    world_.resize(islands_.size());
    for (size_t i = 0; i < islands_.size(); i++) {
      const Island& is = islands_[i];
      const float h = is.transform.rotation * 0.5f;
      const float s = std::sin(h), c = std::cos(h), z = 0.0f;
      const float b2 = (z * z + z * z) + s * s, k1 = c * c - b2, k3 = c * 2.0f;
      const size_t nv = is.mesh.vertices.size() / 3;
      world_[i].resize(nv);
      for (size_t v = 0; v < nv; v++) {
        const float vx = is.mesh.vertices[3 * v], vy = is.mesh.vertices[3 * v + 1], vz = is.mesh.vertices[3 * v + 2];
        const float k2 = ((vx * z + vy * z) + vz * s) * 2.0f;
        const V3d bxv{z * vz - vy * s, s * vx - vz * z, z * vy - vx * z};
        const V3d out{(vx * k1 + z * k2) + bxv.x * k3, (vy * k1 + z * k2) + bxv.y * k3, (vz * k1 + s * k2) + bxv.z * k3};
        world_[i][v] = {out.x + is.transform.translation[0], out.y + is.transform.translation[1],
                        out.z + is.transform.translation[2]};
      }
    }
  }

  // world (left, right) of boundary edge k of island i: left = vertex[edge + 1], right = vertex[edge] (nav_mesh.rs:603-609)
  void edge_points(size_t i, size_t k, V3d& left, V3d& right) const {
    const ValidNavigationMesh& m = islands_[i].mesh;
    const uint32_t p = m.boundary_edges[k].first, e = m.boundary_edges[k].second;
    const uint32_t b = m.poly_edge_begin[p], n = m.poly_edge_begin[p + 1] - b;
    left = world_[i][m.poly_verts[b + (e == n - 1 ? 0 : e + 1)]];
    right = world_[i][m.poly_verts[b + e]];
  }

  // geometry.rs:3-55
  static bool edge_intersection(V3d e1a, V3d e1b, V3d e2a, V3d e2b, float dist_sq, V3d& p0, V3d& p1) {
    const V3d d1 = sub(e1b, e1a);
    const float l1 = dot3(d1, d1);
    const float t2a = clamp01(dot3(d1, sub(e2a, e1a)) / l1), t2b = clamp01(dot3(d1, sub(e2b, e1a)) / l1);
    bool ok = !(t2a <= t2b);
    const V3d d2 = sub(e2b, e2a);
    const float l2 = dot3(d2, d2);
    const float t1a = clamp01(dot3(d2, sub(e1a, e2a)) / l2), t1b = clamp01(dot3(d2, sub(e1b, e2a)) / l2);
    ok = ok && !(t1a <= t1b);
    const V3d p2on1_0 = add(mul(d1, t2a), e1a), p2on1_1 = add(mul(d1, t2b), e1a);
    const V3d p1on2_0 = add(mul(d2, t1a), e2a), p1on2_1 = add(mul(d2, t1b), e2a);
    const V3d da = sub(p2on1_1, p1on2_0), db = sub(p2on1_0, p1on2_1);
    ok = ok && !((dot3(da, da) > dist_sq) || (dot3(db, db) > dist_sq));
    p0 = mul(add(p2on1_1, p1on2_0), 0.5f);
    p1 = mul(add(p2on1_0, p1on2_1), 0.5f);
    return ok;
  }

  // link_edges_between_islands (nav_data.rs:1254-1328) for every pair of islands whose bounds come close
  void link_islands(float d) {
    const float d2 = d * d;
    const size_t n = islands_.size();
    std::vector<std::array<float, 6>> bounds(n);
    std::vector<bool> has(n, false);
    for (size_t i = 0; i < n; i++) {
      if (world_[i].empty()) continue;
      has[i] = true;
      bounds[i] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (const V3d& v : world_[i]) {
        const float c[3] = {v.x, v.y, v.z};
        for (int k = 0; k < 3; k++) bounds[i][k] = std::fmin(bounds[i][k], c[k]), bounds[i][3 + k] = std::fmax(bounds[i][3 + k], c[k]);
      }
    }
    auto comp = [](const V3d& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); };
    std::vector<Link> created;
    for (size_t i1 = 0; i1 < n; i1++)
      for (size_t i2 = i1 + 1; i2 < n; i2++) {
        if (!has[i1] || !has[i2]) continue;
        const ValidNavigationMesh &m1 = islands_[i1].mesh, &m2 = islands_[i2].mesh;
        if (m1.boundary_edges.empty() || m2.boundary_edges.empty()) continue;
        float mn1[3], mx1[3];
        bool overlap = true;
        for (int k = 0; k < 3; k++) {
          mn1[k] = bounds[i1][k] - d, mx1[k] = bounds[i1][3 + k] + d;
          overlap = overlap && mn1[k] <= bounds[i2][3 + k] && bounds[i2][k] <= mx1[k];
        }
        if (!overlap) continue;
        // edges of each island near the other one
        struct E {
          size_t k;
          V3d l, r;
        };
        std::vector<E> near1, near2;
        for (size_t k = 0; k < m1.boundary_edges.size(); k++) {
          E e{k, {}, {}};
          edge_points(i1, k, e.l, e.r);
          bool in = true;
          for (int c = 0; c < 3; c++)
            in = in && std::fmin(comp(e.l, c), comp(e.r, c)) <= bounds[i2][3 + c] + d &&
                 std::fmax(comp(e.l, c), comp(e.r, c)) >= bounds[i2][c] - d;
          if (in) near1.push_back(e);
        }
        for (size_t k = 0; k < m2.boundary_edges.size(); k++) {
          E e{k, {}, {}};
          edge_points(i2, k, e.l, e.r);
          bool in = true;
          for (int c = 0; c < 3; c++)
            in = in && std::fmin(comp(e.l, c), comp(e.r, c)) <= mx1[c] && std::fmax(comp(e.l, c), comp(e.r, c)) >= mn1[c];
          if (in) near2.push_back(e);
        }
        // island_2 edge outer loop, island_1 edge inner: the island_2 edge's box is the one expanded by d
        for (const E& e2 : near2)
          for (const E& e1 : near1) {
            bool ov = true;
            for (int c = 0; c < 3; c++) {
              const float a_min = std::fmin(comp(e1.l, c), comp(e1.r, c)), a_max = std::fmax(comp(e1.l, c), comp(e1.r, c));
              const float b_min = std::fmin(comp(e2.l, c), comp(e2.r, c)) - d, b_max = std::fmax(comp(e2.l, c), comp(e2.r, c)) + d;
              ov = ov && a_min <= b_max && b_min <= a_max;
            }
            if (!ov) continue;
            V3d p0, p1;
            if (!edge_intersection(e1.l, e1.r, e2.l, e2.r, d2, p0, p1)) continue;
            const V3d dp = sub(p0, p1);
            if (dot3(dp, dp) < d2) continue;
            const uint32_t poly1 = m1.boundary_edges[e1.k].first, poly2 = m2.boundary_edges[e2.k].first;
            const size_t at = created.size();
            created.push_back({(uint32_t)i1, poly1, (uint32_t)i2, poly2, m2.poly_type[poly2], p0, p1, at + 1});
            created.push_back({(uint32_t)i2, poly2, (uint32_t)i1, poly1, m1.poly_type[poly1], p1, p0, at});
          }
      }
    // canonical order: by source node, then creation order
    std::vector<size_t> order(created.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      if (created[a].src_island != created[b].src_island) return created[a].src_island < created[b].src_island;
      return created[a].src_polygon < created[b].src_polygon;
    });
    std::vector<size_t> index_of(created.size());
    for (size_t i = 0; i < order.size(); i++) index_of[order[i]] = i;
    links_.clear();
    for (size_t i = 0; i < order.size(); i++) {
      Link l = created[order[i]];
      l.pair = index_of[l.pair];
      links_.push_back(l);
    }
  }

  // update_modified_node (nav_data.rs:827-1011); the `geo` clip of the reference restated, as in the Python
  // mirror, as interval subtraction along each boundary edge (f64 like the reference's geo types)
  void modified_nodes(float d_f) {
    std::map<std::pair<uint32_t, uint32_t>, std::vector<size_t>> by_node;  // ascending (island, polygon)
    for (size_t i = 0; i < links_.size(); i++) by_node[{links_[i].src_island, links_[i].src_polygon}].push_back(i);
    const double tol = std::max((double)d_f * 4.0, 1e-5);
    for (const auto& kv : by_node) {
      const uint32_t ii = kv.first.first, poly = kv.first.second;
      const ValidNavigationMesh& m = islands_[ii].mesh;
      const uint32_t b = m.poly_edge_begin[poly], n = m.poly_edge_begin[poly + 1] - b;
      const uint32_t nverts = (uint32_t)(m.vertices.size() / 3);
      Modified node{ii, poly, {}, {}};
      bool have_center = false;
      double cx = 0.0, cy = 0.0;
      for (uint32_t e = 0; e < n; e++) {
        if (m.edge_neighbor[b + e] != LMB_NONE) continue;
        const uint32_t ia = m.poly_verts[b + e], ib = m.poly_verts[b + (e + 1) % n];
        const double ax = world_[ii][ia].x, ay = world_[ii][ia].y, bx = world_[ii][ib].x, by = world_[ii][ib].y;
        const double abx = bx - ax, aby = by - ay;
        const double L2 = abx * abx + aby * aby;
        if (L2 == 0.0) continue;
        const double L = std::sqrt(L2);
        std::vector<std::pair<double, double>> cuts;
        for (size_t li : kv.second) {
          const Link& l = links_[li];
          const double p0x = l.p0.x, p0y = l.p0.y, p1x = l.p1.x, p1y = l.p1.y;
          const double dist0 = std::fabs(abx * (p0y - ay) - aby * (p0x - ax)) / L;
          const double dist1 = std::fabs(abx * (p1y - ay) - aby * (p1x - ax)) / L;
          if (dist0 > tol || dist1 > tol) continue;
          const double t0 = ((p0x - ax) * abx + (p0y - ay) * aby) / L2, t1 = ((p1x - ax) * abx + (p1y - ay) * aby) / L2;
          const double lo = std::max(std::min(t0, t1), 0.0), hi = std::min(std::max(t0, t1), 1.0);
          if (hi - lo <= 1e-9) continue;
          cuts.emplace_back(lo, hi);
        }
        std::sort(cuts.begin(), cuts.end());
        std::vector<std::pair<double, double>> merged, pieces;
        for (const auto& c : cuts) {
          if (!merged.empty() && c.first <= merged.back().second + 1e-9)
            merged.back().second = std::max(merged.back().second, c.second);
          else
            merged.push_back(c);
        }
        double cur = 0.0;
        for (const auto& c : merged) {
          if (c.first > cur) pieces.emplace_back(cur, c.first);
          cur = std::max(cur, c.second);
        }
        if (cur < 1.0) pieces.emplace_back(cur, 1.0);
        for (const auto& pc : pieces) {
          const double sx = ax + abx * pc.first, sy = ay + aby * pc.first, ex = ax + abx * pc.second, ey = ay + aby * pc.second;
          auto snap = [&](double px, double py) -> int64_t {  // nearest polygon vertex within squared distance 0.01
            int64_t best = -1;
            double best_d = 0.0;
            for (uint32_t k = 0; k < n; k++) {
              const uint32_t vi = m.poly_verts[b + k];
              const double dx = (double)world_[ii][vi].x - px, dy = (double)world_[ii][vi].y - py;
              const double dd = dx * dx + dy * dy;
              if (best < 0 || dd < best_d) best = vi, best_d = dd;
            }
            return best >= 0 && best_d < 0.01 ? best : -1;
          };
          int64_t si = snap(sx, sy), ei = snap(ex, ey);
          if (si >= 0 && ei >= 0 && si == ei) continue;
          if (si < 0) {
            node.new_vertices.emplace_back((float)sx, (float)sy);
            si = nverts + (int64_t)node.new_vertices.size() - 1;
          }
          if (ei < 0) {
            node.new_vertices.emplace_back((float)ex, (float)ey);
            ei = nverts + (int64_t)node.new_vertices.size() - 1;
          }
          if (!have_center) {
            const V3d c = apply_point(ii, {m.poly_center[3 * poly], m.poly_center[3 * poly + 1], m.poly_center[3 * poly + 2]});
            cx = c.x, cy = c.y;
            have_center = true;
          }
          const double scx = sx - cx, scy = sy - cy, ecx = ex - cx, ecy = ey - cy;
          if (scx * ecy - scy * ecx < 0.0) std::swap(si, ei);
          node.new_boundary.emplace_back((uint32_t)si, (uint32_t)ei);
        }
      }
      modified_.push_back(std::move(node));
    }
  }

  // Transform::apply of one point, scalar form of build_world (nav_data.transform_apply in the Python mirror)
  V3d apply_point(size_t i, V3d v) const {
    const Island& is = islands_[i];
    const float h = is.transform.rotation * 0.5f;
    const float s = std::sin(h), c = std::cos(h), z = 0.0f;
    const float b2 = (z * z + z * z) + s * s, k1 = c * c - b2, k3 = c * 2.0f;
    const float k2 = ((v.x * z + v.y * z) + v.z * s) * 2.0f;
    const V3d bxv{z * v.z - v.y * s, s * v.x - v.z * z, z * v.y - v.x * z};
    const V3d out{(v.x * k1 + z * k2) + bxv.x * k3, (v.y * k1 + z * k2) + bxv.y * k3, (v.z * k1 + s * k2) + bxv.z * k3};
    return {out.x + is.transform.translation[0], out.y + is.transform.translation[1], out.z + is.transform.translation[2]};
  }

  // update_regions (nav_data.rs:1089-1156): numbers in order of first use by a link, union by smaller root
  void update_regions() {
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> number;
    std::vector<uint32_t> parent;
    auto find = [&](uint32_t x) {
      while (parent[x] != x) x = parent[x] = parent[parent[x]];
      return x;
    };
    auto num = [&](uint32_t island, uint32_t polygon) {
      const std::pair<uint32_t, uint32_t> rid{island, islands_[island].mesh.poly_region[polygon]};
      auto it = number.find(rid);
      if (it != number.end()) return it->second;
      const uint32_t v = (uint32_t)parent.size();
      number[rid] = v;
      parent.push_back(v);
      return v;
    };
    for (const Link& l : links_) {
      const uint32_t a = num(l.src_island, l.src_polygon), b = num(l.dst_island, l.dst_polygon);
      const uint32_t ra = find(a), rb = find(b);
      if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
    }
    region_root_.resize(parent.size());
    for (uint32_t i = 0; i < parent.size(); i++) region_root_[i] = find(i);
    region_number_of_ = std::move(number);
  }

  void flatten() {
    const size_t n = islands_.size();
    poly_begin_.assign(n + 1, 0u), vert_begin_.assign(n + 1, 0u);
    for (size_t i = 0; i < n; i++) {
      const ValidNavigationMesh& m = islands_[i].mesh;
      const uint32_t pb = poly_begin_[i], vb = vert_begin_[i];
      for (int k = 0; k < 3; k++) translation_.push_back(islands_[i].transform.translation[k]);
      rotation_.push_back(islands_[i].transform.rotation);
      const std::array<float, 6> empty{1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f};
      const std::array<float, 6>& bd = m.bounds_empty ? empty : m.mesh_bounds;
      bounds_.insert(bounds_.end(), bd.begin(), bd.end());
      vertices_.insert(vertices_.end(), m.vertices.begin(), m.vertices.end());
      for (size_t p = 0; p < m.n_polys(); p++) edge_begin_.push_back((uint32_t)edge_vertex_.size() + m.poly_edge_begin[p]);
      for (uint32_t v : m.poly_verts) edge_vertex_.push_back(v + vb);
      for (uint32_t nb : m.edge_neighbor) edge_neighbor_.push_back(nb == LMB_NONE ? LMB_NONE : nb + pb);
      edge_reverse_.insert(edge_reverse_.end(), m.edge_reverse.begin(), m.edge_reverse.end());
      poly_type_.insert(poly_type_.end(), m.poly_type.begin(), m.poly_type.end());
      poly_region_.insert(poly_region_.end(), m.poly_region.begin(), m.poly_region.end());
      poly_bounds_.insert(poly_bounds_.end(), m.poly_bounds.begin(), m.poly_bounds.end());
      poly_center_.insert(poly_center_.end(), m.poly_center.begin(), m.poly_center.end());
      // height meshes: base vertex / triangle become global; islands without one get empty height polygons
      if (m.has_height) {
        any_height_ = true;
        const uint32_t hvb = (uint32_t)(hverts_.size() / 3), htb = (uint32_t)(htris_.size() / 3);
        for (size_t p = 0; p < m.n_polys(); p++) {
          hpoly_.push_back(m.hpoly[4 * p] + hvb), hpoly_.push_back(m.hpoly[4 * p + 1]);
          hpoly_.push_back(m.hpoly[4 * p + 2] + htb), hpoly_.push_back(m.hpoly[4 * p + 3]);
        }
        hverts_.insert(hverts_.end(), m.hverts.begin(), m.hverts.end());
        htris_.insert(htris_.end(), m.htris.begin(), m.htris.end());
      } else {
        hpoly_.insert(hpoly_.end(), 4 * m.n_polys(), 0u);
      }
      has_height_.push_back(m.has_height ? 1 : 0);
      poly_begin_[i + 1] = pb + (uint32_t)m.n_polys();
      vert_begin_[i + 1] = vb + (uint32_t)(m.vertices.size() / 3);
    }
    // poly_edge_begin: per-island begins were pushed relative to the running edge count BEFORE the island's
    // edges were appended, so they are already global; close the CSR
    edge_begin_.push_back((uint32_t)edge_vertex_.size());
    const uint32_t P = poly_begin_[n];
    // links
    const uint32_t nl = (uint32_t)links_.size();
    node_link_begin_.assign((size_t)P + 1, 0u);
    for (const Link& l : links_) {
      const uint32_t src = node_id(l.src_island, l.src_polygon);
      link_src_.push_back(src);
      link_dst_.push_back(node_id(l.dst_island, l.dst_polygon));
      link_dst_type_.push_back(l.dst_type);
      const float pt[6] = {l.p0.x, l.p0.y, l.p0.z, l.p1.x, l.p1.y, l.p1.z};
      link_portal_.insert(link_portal_.end(), pt, pt + 6);
      link_kind_.push_back(LMB_LINK_BOUNDARY);
      link_reverse_.push_back((uint32_t)l.pair);
      link_cost_.push_back(0.0f), link_anim_kind_.push_back(0u), link_anim_id_.push_back(0u);
      node_link_begin_[src + 1]++;
    }
    for (uint32_t p = 0; p < P; p++) node_link_begin_[p + 1] += node_link_begin_[p];
    node_links_.resize(nl);  // links are sorted by source node already: ascending link index within a node
    for (uint32_t l = 0; l < nl; l++) node_links_[l] = l;
    // modified nodes (by_node iterates ascending (island, polygon) = ascending node id)
    mod_edge_begin_.push_back(0u), mod_vert_begin_.push_back(0u);
    for (const Modified& m : modified_) {
      mod_node_.push_back(node_id(m.island, m.polygon));
      for (const auto& e : m.new_boundary) mod_edges_.push_back(e.first), mod_edges_.push_back(e.second);
      for (const auto& v : m.new_vertices) mod_verts_.push_back(v.first), mod_verts_.push_back(v.second);
      mod_edge_begin_.push_back((uint32_t)(mod_edges_.size() / 2));
      mod_vert_begin_.push_back((uint32_t)(mod_verts_.size() / 2));
    }
    // regions
    node_region_number_.assign(P, LMB_NONE);
    for (const auto& kv : region_number_of_) {
      const uint32_t i = kv.first.first, region = kv.first.second;
      for (uint32_t p = poly_begin_[i]; p < poly_begin_[i + 1]; p++)
        if (poly_region_[p] == region) node_region_number_[p] = kv.second;
    }

    lmb_navdata_desc& d = desc_;
    d = lmb_navdata_desc{};
    d.struct_size = (uint32_t)sizeof(lmb_navdata_desc);
    d.n_islands = (uint32_t)n;
    d.island_translation = translation_.data(), d.island_rotation = rotation_.data(), d.island_bounds = bounds_.data();
    d.island_poly_begin = poly_begin_.data(), d.island_vert_begin = vert_begin_.data();
    d.n_vertices = vert_begin_[n], d.vertices = vertices_.data();
    d.n_polys = P, d.poly_edge_begin = edge_begin_.data(), d.n_edges = (uint32_t)edge_vertex_.size();
    d.edge_vertex = edge_vertex_.data(), d.edge_neighbor = edge_neighbor_.data(), d.edge_reverse = edge_reverse_.data();
    d.poly_type = poly_type_.data(), d.poly_region = poly_region_.data();
    d.poly_bounds = poly_bounds_.data(), d.poly_center = poly_center_.data();
    d.n_type_costs = (uint32_t)cost_index_.size();
    d.type_cost_index = cost_index_.data(), d.type_cost_value = cost_value_.data();
    if (any_height_) {
      d.island_has_height = has_height_.data(), d.hpoly = hpoly_.data();
      d.n_hverts = (uint32_t)(hverts_.size() / 3), d.hverts = hverts_.data();
      d.n_htris = (uint32_t)(htris_.size() / 3), d.htris = htris_.data();
    }
    d.n_links = nl;
    d.link_dst_node = link_dst_.data(), d.link_dst_type = link_dst_type_.data(), d.link_portal = link_portal_.data();
    d.link_kind = link_kind_.data(), d.link_reverse = link_reverse_.data(), d.link_dst_portal = link_portal_.data();
    d.link_cost = link_cost_.data(), d.link_anim_kind = link_anim_kind_.data(), d.link_anim_id = link_anim_id_.data();
    d.node_link_begin = node_link_begin_.data(), d.node_links = node_links_.data();
    d.n_modified = (uint32_t)modified_.size();
    d.modified_node = mod_node_.data(), d.modified_edge_begin = mod_edge_begin_.data(), d.modified_edges = mod_edges_.data();
    d.modified_vert_begin = mod_vert_begin_.data(), d.modified_verts = mod_verts_.data();
    d.node_region_number = node_region_number_.data();
    d.n_region_numbers = (uint32_t)region_root_.size(), d.region_root = region_root_.data();
  }

  std::vector<Island> islands_;
  std::vector<std::vector<V3d>> world_;
  std::vector<Link> links_;
  std::vector<Modified> modified_;
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> region_number_of_;
  std::vector<uint32_t> region_root_;
  // flattened arrays
  std::vector<float> translation_, rotation_, bounds_, vertices_, poly_bounds_, poly_center_, cost_value_, link_portal_,
      link_cost_, mod_verts_;
  std::vector<uint32_t> poly_begin_, vert_begin_, edge_begin_, edge_vertex_, edge_neighbor_, edge_reverse_, poly_type_,
      poly_region_, cost_index_, link_src_, link_dst_, link_dst_type_, link_kind_, link_reverse_, link_anim_kind_,
      link_anim_id_, node_link_begin_, node_links_, mod_node_, mod_edge_begin_, mod_edges_, mod_vert_begin_,
      node_region_number_;
  std::vector<uint32_t> hpoly_;
  std::vector<float> hverts_;
  std::vector<uint8_t> htris_, has_height_;
  bool any_height_ = false;
  lmb_navdata_desc desc_{};
};

using IslandId = Key;

// `Archipelago` with its islands (lib.rs:139-175, 270-281): islands live in a DenseSlotMap, a change marks the
// navigation data dirty, and `update` re-links and re-uploads it first — through lmb_navdata_update with the map
// from the previous flattening's islands / links to the new one, so that exactly the paths `Path::is_valid`
// keeps survive (the islands deleted or touched since the last update are the invalidated ones).
template <class Backend>
class BasicIslandArchipelago : public BasicArchipelago<Backend> {
  using Base = BasicArchipelago<Backend>;

 public:
  BasicIslandArchipelago(typename Backend::Ctx* ctx, ArchipelagoOptions options, uint32_t max_agents)
      : Base(ctx, options, max_agents), ctx_(ctx) {}

  IslandId add_island(Island island) {  // lib.rs:139-146
    dirty_ = true;
    return islands_.insert(std::move(island));
  }
  void remove_island(IslandId id) {  // lib.rs:161-167
    if (!islands_.remove(id)) throw std::out_of_range("Island should be present in the Archipelago");
    dirty_ = true;
    changed_.push_back(id);
  }
  const Island* get_island(IslandId id) const { return islands_.get(id); }
  // IslandMut (nav_data.rs:1193-1223): handing out mutable access marks the island dirty
  Island* get_island_mut(IslandId id) {
    Island* is = islands_.get(id);
    if (is) {
      dirty_ = true;
      changed_.push_back(id);
    }
    return is;
  }
  const std::vector<IslandId>& get_island_ids() const { return islands_.keys(); }
  bool set_type_index_cost(uint32_t type_index, float cost) {  // nav_data.rs:229-241: SetTypeIndexCostError on cost <= 0
    if (!(cost > 0.0f)) return false;
    for (auto& kv : costs_)
      if (kv.first == type_index) {
        kv.second = cost;
        costs_dirty_ = true;
        return true;
      }
    costs_.emplace_back(type_index, cost);
    costs_dirty_ = true;
    return true;
  }
  const NavigationData* navigation_data() const { return nav_.get(); }

  void update(float delta_time) {
    sync_navigation_data();
    Base::update(delta_time);
  }

  void sync_navigation_data() {
    if (nav_ && !dirty_ && !costs_dirty_) return;
    std::vector<Island> list(islands_.values().begin(), islands_.values().end());
    const std::vector<IslandId> keys = islands_.keys();
    auto nav = std::make_unique<NavigationData>(std::move(list), costs_, 0.01f);  // edge_link_distance, lib.rs:273-276
    if (!nav_) {
      check(Backend::navdata_upload(ctx_, &nav->desc()));
    } else {
      auto is_changed = [&](IslandId k) { return std::find(changed_.begin(), changed_.end(), k) != changed_.end(); };
      auto new_index = [&](IslandId k) -> uint32_t {
        for (size_t i = 0; i < keys.size(); i++)
          if (keys[i] == k) return (uint32_t)i;
        return LMB_NONE;
      };
      std::vector<uint32_t> island_map(keys_.size() + 1, LMB_NONE), link_map(nav_->n_links() + 1, LMB_NONE);
      for (size_t i = 0; i < keys_.size(); i++)
        if (!is_changed(keys_[i])) island_map[i] = new_index(keys_[i]);
      for (size_t l = 0; l < nav_->n_links(); l++) {
        const NavigationData::LinkInfo o = nav_->link_info(l);
        const IslandId ks = keys_[o.src_island], kd = keys_[o.dst_island];
        if (is_changed(ks) || is_changed(kd)) continue;
        const uint32_t ns = new_index(ks), ndst = new_index(kd);
        if (ns == LMB_NONE || ndst == LMB_NONE) continue;
        for (size_t m = 0; m < nav->n_links(); m++) {
          const NavigationData::LinkInfo c = nav->link_info(m);
          if (c.src_island == ns && c.dst_island == ndst && c.src_polygon == o.src_polygon &&
              c.dst_polygon == o.dst_polygon && std::memcmp(c.portal.data(), o.portal.data(), sizeof(float) * 6) == 0) {
            link_map[l] = (uint32_t)m;
            break;
          }
        }
      }
      lmb_nav_remap remap{};
      remap.struct_size = (uint32_t)sizeof(remap);
      remap.n_old_islands = (uint32_t)keys_.size(), remap.island_map = island_map.data();
      remap.n_old_links = (uint32_t)nav_->n_links(), remap.link_map = link_map.data();
      check(Backend::navdata_update(ctx_, &nav->desc(), &remap));
    }
    nav_ = std::move(nav);
    keys_ = keys;
    changed_.clear();
    dirty_ = costs_dirty_ = false;
  }

 private:
  void check(int32_t st) {
    if (st != LMB_OK) throw Error(st, "lmb_navdata_update: " + Backend::last_error(ctx_));
  }
  typename Backend::Ctx* ctx_;
  DenseSlotMap<Island> islands_;
  std::vector<std::pair<uint32_t, float>> costs_;
  std::unique_ptr<NavigationData> nav_;
  std::vector<IslandId> keys_, changed_;  // island order of the uploaded flattening; islands invalidated since
  bool dirty_ = true, costs_dirty_ = false;
};

using IslandArchipelago = BasicIslandArchipelago<CudaBackend>;

}  // namespace landmass_b200
