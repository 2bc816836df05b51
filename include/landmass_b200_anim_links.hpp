// landmass_b200_anim_links.hpp — the geometry under animation-link creation, in C++17: `clip_edge_to_triangle`
// (geometry.rs:63-170), `BoundingBox::intersects_ray_segment` (util.rs:219-262, 491-502) and
// `ValidNavigationMesh::sample_edge` (nav_mesh.rs:819-923), f32 in the reference's operation order, NaN / infinity
// behaviour included.  Same restatement as landmass_b200/anim_links.py, compared with it bit for bit in
// tests/test_host_mirror_cpp.py.  The link assembly on top of these (nav_data.rs:516-825: portals of a link's two
// edges on every island, interval intersection, `PossibleRegionLink`s) is mirrored in Python only so far.
#pragma once
#include <algorithm>
#include <cmath>
#include <optional>
#include <utility>
#include <vector>

#include "landmass_b200.hpp"

namespace landmass_b200 {
namespace anim_detail {
inline float perp_dot(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
inline float dot2(float ax, float ay, float bx, float by) { return ax * bx + ay * by; }
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }  // FloatExt::lerp
// Rust's f32::max / min as the Python mirror spells them (`max(a, b)`: b if b > a else a)
inline float pick_max(float a, float b) { return b > a ? b : a; }
inline float pick_min(float a, float b) { return b < a ? b : a; }
}  // namespace anim_detail

// geometry.rs:63-170: the part (t0, t1) of `edge` that lies over the triangle and within `max_vertical_distance`
// of it; empty when there is none.
inline std::optional<std::pair<float, float>> clip_edge_to_triangle(const Vec3 tri[3], const Vec3 edge[2],
                                                                    float max_vertical_distance) {
  using namespace anim_detail;
  const float mv = max_vertical_distance;
  const float dx[3] = {tri[1][0] - tri[0][0], tri[2][0] - tri[1][0], tri[0][0] - tri[2][0]};
  const float dy[3] = {tri[1][1] - tri[0][1], tri[2][1] - tri[1][1], tri[0][1] - tri[2][1]};
  const float ex = edge[1][0] - edge[0][0], ey = edge[1][1] - edge[0][1];
  const float ox[3] = {tri[0][0] - edge[0][0], tri[1][0] - edge[0][0], tri[2][0] - edge[0][0]};
  const float oy[3] = {tri[0][1] - edge[0][1], tri[1][1] - edge[0][1], tri[2][1] - edge[0][1]};
  const float v0x = dx[0], v0y = dy[0], v1x = dx[2], v1y = dy[2];
  const float denominator = perp_dot(v0x, v0y, v1x, v1y);
  auto estimate_height = [&](float px, float py) {
    const float v2x = px - tri[0][0], v2y = py - tri[0][1];
    const float v = perp_dot(v2x, v2y, v1x, v1y) / denominator;
    const float w = perp_dot(v2x, v2y, v0x, v0y) / denominator;
    const float u = (1.0f - v) - w;
    return (tri[0][2] * u + tri[1][2] * v) + tri[2][2] * w;
  };
  const float h0 = estimate_height(edge[0][0], edge[0][1]), h1 = estimate_height(edge[1][0], edge[1][1]);
  float ts[3];
  for (int k = 0; k < 3; k++) {
    const float det = -perp_dot(ex, ey, dx[k], dy[k]);
    ts[k] = det == 0.0f ? INFINITY : dot2(-dy[k], dx[k], ox[k], oy[k]) / det;
  }
  float tmin = 0.0f, tmax = 1.0f;
  for (int k = 0; k < 3; k++) {
    if (std::isinf(ts[k])) {
      if (perp_dot(ox[k], oy[k], dx[k], dy[k]) < 0.0f) return std::nullopt;
    } else if (perp_dot(ex, ey, dx[k], dy[k]) < 0.0f) {
      tmin = pick_max(tmin, ts[k]);
    } else {
      tmax = pick_min(tmax, ts[k]);
    }
  }
  if (tmin >= 1.0f || tmax <= 0.0f) return std::nullopt;
  const float th_min = lerp(h0, h1, tmin), th_max = lerp(h0, h1, tmax);
  const float eh_min = lerp(edge[0][2], edge[1][2], tmin), eh_max = lerp(edge[0][2], edge[1][2], tmax);
  const float slope = (eh_max - eh_min) - (th_max - th_min);
  if (slope == 0.0f) {
    if (std::fabs(th_min - eh_min) < mv) return std::make_pair(tmin, tmax);
    return std::nullopt;
  }
  const float ta = ((mv - eh_min) + th_min) / slope, tb = ((-mv - eh_min) + th_min) / slope;
  float t0 = pick_min(ta, tb), t1 = pick_max(ta, tb);
  if (t0 >= 1.0f || t1 <= 0.0f) return std::nullopt;
  t0 = pick_max(t0, 0.0f), t1 = pick_min(t1, 1.0f);
  return std::make_pair(lerp(tmin, tmax, t0), lerp(tmin, tmax, t1));
}

// BoundingBox::intersects_ray_segment (util.rs:219-262) with RaySegment::new (util.rs:491-502)
inline bool intersects_ray_segment(const float bmin[3], const float bmax[3], const Vec3& p0, const Vec3& p1) {
  using namespace anim_detail;
  float inv[3];
  int sign[3];
  for (int k = 0; k < 3; k++) {
    const float d = p1[k] - p0[k];
    inv[k] = 1.0f / d;
    sign[k] = d < 0.0f ? 1 : 0;
  }
  auto b = [&](int k, int s) { return s ? bmax[k] : bmin[k]; };
  float tmin = (b(0, sign[0]) - p0[0]) * inv[0], tmax = (b(0, 1 - sign[0]) - p0[0]) * inv[0];
  if (tmin >= 1.0f || tmax <= 0.0f) return false;
  const float tymin = (b(1, sign[1]) - p0[1]) * inv[1], tymax = (b(1, 1 - sign[1]) - p0[1]) * inv[1];
  if (tmin > tymax || tymin > tmax) return false;
  tmin = pick_max(tmin, tymin), tmax = pick_min(tmax, tymax);
  if (std::isnan(tmin)) tmin = tymin;
  if (tmin >= 1.0f || tmax <= 0.0f) return false;
  const float tzmin = (b(2, sign[2]) - p0[2]) * inv[2], tzmax = (b(2, 1 - sign[2]) - p0[2]) * inv[2];
  if (tmin > tzmax || tzmin > tmax) return false;
  tmin = std::fmax(tmin, tzmin), tmax = std::fmin(tmax, tzmax);  // f32::max / min: NaN ignored
  return tmin < 1.0f && tmax > 0.0f;
}

struct SampledEdgePortion {
  uint32_t node;
  float t0, t1;
};

// ValidNavigationMesh::sample_edge (nav_mesh.rs:819-923): for every node the merged intervals of `edge` (island-local
// coordinates) that lie on it, nodes ascending (the reference's bounding-box hierarchy only prunes)
inline std::vector<SampledEdgePortion> sample_edge(const ValidNavigationMesh& mesh, const Vec3 edge[2],
                                                   float max_vertical_distance) {
  const float mv = max_vertical_distance;
  std::vector<SampledEdgePortion> out;
  auto hv = [&](size_t i) { return Vec3{mesh.hverts[3 * i], mesh.hverts[3 * i + 1], mesh.hverts[3 * i + 2]}; };
  auto mvx = [&](size_t i) { return Vec3{mesh.vertices[3 * i], mesh.vertices[3 * i + 1], mesh.vertices[3 * i + 2]}; };
  for (uint32_t node = 0; node < mesh.n_polys(); node++) {
    const float* b = &mesh.poly_bounds[6 * (size_t)node];
    if (!(b[0] <= b[3])) continue;
    const float bmin[3] = {b[0], b[1], b[2] - mv}, bmax[3] = {b[3], b[4], b[5] + mv};
    if (!(bmax[0] - bmin[0] >= 0.0f && bmax[1] - bmin[1] >= 0.0f && bmax[2] - bmin[2] >= 0.0f)) continue;
    if (!intersects_ray_segment(bmin, bmax, edge[0], edge[1])) continue;
    std::vector<std::pair<float, float>> intervals;
    auto add = [&](const Vec3& a, const Vec3& bb, const Vec3& c) {
      const Vec3 tri[3] = {a, bb, c};
      if (auto iv = clip_edge_to_triangle(tri, edge, mv)) intervals.push_back(*iv);
    };
    if (mesh.has_height) {
      const uint32_t bv = mesh.hpoly[4 * (size_t)node], bt = mesh.hpoly[4 * (size_t)node + 2], nt = mesh.hpoly[4 * (size_t)node + 3];
      for (uint32_t t = bt; t < bt + nt; t++)
        add(hv(bv + mesh.htris[3 * (size_t)t]), hv(bv + mesh.htris[3 * (size_t)t + 1]), hv(bv + mesh.htris[3 * (size_t)t + 2]));
    } else {
      const uint32_t pb = mesh.poly_edge_begin[node], n = mesh.poly_edge_begin[node + 1] - pb;
      for (uint32_t i = 2; i < n; i++) add(mvx(mesh.poly_verts[pb]), mvx(mesh.poly_verts[pb + i - 1]), mvx(mesh.poly_verts[pb + i]));
    }
    if (intervals.empty()) continue;
    // FloatOrd: NaN sorts last; stable
    std::stable_sort(intervals.begin(), intervals.end(), [](const auto& x, const auto& y) {
      const bool nx = std::isnan(x.first), ny = std::isnan(y.first);
      if (nx != ny) return ny;
      return x.first < y.first;
    });
    std::pair<float, float> cur = intervals[0];
    for (size_t k = 1; k < intervals.size(); k++) {
      if (intervals[k].first - cur.second <= 1e-5f) {
        cur.second = intervals[k].second;
      } else {
        out.push_back({node, cur.first, cur.second});
        cur = intervals[k];
      }
    }
    out.push_back({node, cur.first, cur.second});
  }
  return out;
}

}  // namespace landmass_b200
