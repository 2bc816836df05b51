// Kernel 1 — batched point -> node sampling.
//
// Replaces `NavigationData::sample_point` (crates/landmass/src/nav_data.rs:269-324)
// and `ValidNavigationMesh::sample_point` (nav_mesh.rs:614-737).  The reference scans
// every polygon of every island per query (O(P)); here each island carries a uniform
// grid over its polygon AABBs and a query walks only the cells its sample box touches.
// Every candidate is re-tested with the reference's exact closed-interval box test and
// scored with op-for-op identical f32 arithmetic; the winner is the lexicographic
// minimum of (score, polygon index, triangle index), which is what the reference's
// in-order scan with strict `<` replacement produces.
//
// One thread per query: a query touches 1-4 cells and a handful of polygons, the data
// (<= ~60 MB) is L2-resident, so the kernel is latency-bound; occupancy hides it.
#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {

// nav_mesh.rs:634-665
__device__ __forceinline__ V3 project_to_triangle(V3 t0, V3 t1, V3 t2, V3 point) {
  V3 d0 = t1 - t0, d1 = t2 - t1, d2 = t0 - t2;
  V2 f0 = xy(d0), f1 = xy(d1), f2 = xy(d2);
  V2 p = xy(point);
  if (perp_dot(f0, p - xy(t0)) < 0.0f) {
    float s = fdiv(dot(f0, p - xy(t0)), length_squared(f0));
    return d0 * clampf(s, 0.0f, 1.0f) + t0;
  }
  if (perp_dot(f1, p - xy(t1)) < 0.0f) {
    float s = fdiv(dot(f1, p - xy(t1)), length_squared(f1));
    return d1 * clampf(s, 0.0f, 1.0f) + t1;
  }
  if (perp_dot(f2, p - xy(t2)) < 0.0f) {
    float s = fdiv(dot(f2, p - xy(t2)), length_squared(f2));
    return d2 * clampf(s, 0.0f, 1.0f) + t2;
  }
  V3 normal = -normalize(cross(d0, d2));
  float height = fdiv(dot(normal, point - t0), normal.z);
  return {point.x, point.y, fsub(point.z, height)};
}

struct Best {
  bool have;
  float score;
  uint32_t poly, tri;
  V3 point;
};

__device__ __forceinline__ void test_triangle(V3 t0, V3 t1, V3 t2, V3 point, const lmb_options& o,
                                              uint32_t poly, uint32_t tri, Best& best) {
  V3 q = project_to_triangle(t0, t1, t2, point);
  float dh = distance(xy(point), xy(q));
  float dv = fsub(q.z, point.z);
  if (dh <= o.horizontal_distance && (-o.distance_below <= dv && dv <= o.distance_above)) {
    float d = fadd(fmul(dh, o.vertical_preference_ratio), fabsf(dv));
    bool replace = !best.have || d < best.score ||
                   (d == best.score && (poly < best.poly || (poly == best.poly && tri < best.tri)));
    if (replace) {
      best.have = true;
      best.score = d;
      best.poly = poly;
      best.tri = tri;
      best.point = q;
    }
  }
}

__device__ bool sample_island(const DevNav& nav, const DevIsland& is, V3 point, const lmb_options& o,
                              V3& out_point, uint32_t& out_node) {
  V3 smin = point + V3{-o.horizontal_distance, -o.horizontal_distance, -o.distance_below};
  V3 smax = point + V3{o.horizontal_distance, o.horizontal_distance, o.distance_above};
  int32_t cx0 = grid_cell(smin.x, is.gx0, is.ginv, is.gnx), cx1 = grid_cell(smax.x, is.gx0, is.ginv, is.gnx);
  int32_t cy0 = grid_cell(smin.y, is.gy0, is.ginv, is.gny), cy1 = grid_cell(smax.y, is.gy0, is.ginv, is.gny);
  Best best;
  best.have = false;
  best.score = 0.0f;
  best.poly = best.tri = 0;
  best.point = {0, 0, 0};
  const uint32_t* cell_start = nav.cell_start + is.cell_begin;
  for (int32_t cy = cy0; cy <= cy1; cy++) {
    for (int32_t cx = cx0; cx <= cx1; cx++) {
      uint32_t c = (uint32_t)cy * is.gnx + (uint32_t)cx;
      uint32_t kb = __ldg(cell_start + c), ke = __ldg(cell_start + c + 1);
      for (uint32_t k = kb; k < ke; k++) {
        uint32_t p = __ldg(nav.cell_polys + k);
        const float* b = nav.poly_bounds + (size_t)p * 6;
        float b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2), b3 = __ldg(b + 3), b4 = __ldg(b + 4),
              b5 = __ldg(b + 5);
        if (!(b0 <= b3)) continue;  // Empty bounds never intersect
        // visit a polygon only from the first query cell it overlaps (perf only: the
        // lexicographic-min update is idempotent)
        int32_t px0 = grid_cell(b0, is.gx0, is.ginv, is.gnx), py0 = grid_cell(b1, is.gy0, is.ginv, is.gny);
        if ((px0 > cx0 ? px0 : cx0) != cx || (py0 > cy0 ? py0 : cy0) != cy) continue;
        // BoundingBox::intersects_bounds (util.rs:201-217)
        if (!(smin.x <= b3 && b0 <= smax.x && smin.y <= b4 && b1 <= smax.y && smin.z <= b5 && b2 <= smax.z))
          continue;
        if (is.has_height) {
          const uint32_t* hp = nav.hpoly + (size_t)p * 4;
          uint32_t bv = __ldg(hp), bt = __ldg(hp + 2), nt = __ldg(hp + 3);
          for (uint32_t t = 0; t < nt; t++) {
            const uint8_t* tri = nav.htris + (size_t)(bt + t) * 3;
            test_triangle(ld3(nav.hverts, bv + tri[0]), ld3(nav.hverts, bv + tri[1]),
                          ld3(nav.hverts, bv + tri[2]), point, o, p, t, best);
          }
        } else {
          uint32_t eb = __ldg(nav.poly_edge_begin + p), n = __ldg(nav.poly_edge_begin + p + 1) - eb;
          V3 v0 = ld3(nav.verts_local, __ldg(nav.edge_vertex + eb));
          V3 vp = ld3(nav.verts_local, __ldg(nav.edge_vertex + eb + 1));
          for (uint32_t i = 2; i < n; i++) {
            V3 vi = ld3(nav.verts_local, __ldg(nav.edge_vertex + eb + i));
            test_triangle(v0, vp, vi, point, o, p, i, best);
            vp = vi;
          }
        }
      }
    }
  }
  out_point = best.point;
  out_node = best.poly;
  return best.have;
}

// nav_data.rs:269-324
__device__ bool sample_point(const DevNav& nav, V3 point, const lmb_options& o, V3& out_point,
                             uint32_t& out_node) {
  bool have = false;
  float best_distance = 0.0f;
  for (uint32_t i = 0; i < nav.n_islands; i++) {
    const DevIsland& is = nav.islands[i];
    if (is.bounds_empty) continue;
    V3 rel = island_apply_inverse(is, point);
    V3 mn = V3{is.bmin[0], is.bmin[1], is.bmin[2]} +
            V3{-o.horizontal_distance, -o.horizontal_distance, -o.distance_above};
    V3 mx = V3{is.bmax[0], is.bmax[1], is.bmax[2]} +
            V3{o.horizontal_distance, o.horizontal_distance, o.distance_below};
    V3 size = mx - mn;
    if (!(size.x >= 0.0f && size.y >= 0.0f && size.z >= 0.0f)) continue;
    if (!(mn.x <= rel.x && rel.x <= mx.x && mn.y <= rel.y && rel.y <= mx.y && mn.z <= rel.z && rel.z <= mx.z))
      continue;
    V3 sp;
    uint32_t sn;
    if (!sample_island(nav, is, rel, o, sp, sn)) continue;
    float d = distance_squared(rel, sp);
    if (have && d >= best_distance) continue;
    have = true;
    best_distance = d;
    out_point = island_apply(is, sp);
    out_node = sn;
  }
  return have;
}

// points: [n*3]; mask (optional): skip queries whose mask word & mask_bits != mask_want.
__global__ void __launch_bounds__(128)
k_sample_points(DevNav nav, lmb_options o, uint32_t n, const float* __restrict__ points,
                const uint32_t* __restrict__ flags, uint32_t skip_if_any, uint32_t require_all,
                float* __restrict__ out_points, uint32_t* __restrict__ out_nodes) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t node = LMB_NONE;
  V3 q{0.0f, 0.0f, 0.0f};
  bool active = true;
  if (flags) {
    uint32_t f = flags[i];
    active = (f & skip_if_any) == 0 && (f & require_all) == require_all;
  }
  if (active) {
    V3 p{points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]};
    if (!sample_point(nav, p, o, q, node)) {
      node = LMB_NONE;
      q = {0.0f, 0.0f, 0.0f};
    }
  }
  out_points[3 * (size_t)i] = q.x;
  out_points[3 * (size_t)i + 1] = q.y;
  out_points[3 * (size_t)i + 2] = q.z;
  out_nodes[i] = node;
}

void launch_sample_points(const DevNav& nav, const lmb_options& o, uint32_t n, const float* points,
                          const uint32_t* flags, uint32_t skip_if_any, uint32_t require_all,
                          float* out_points, uint32_t* out_nodes, cudaStream_t stream) {
  if (n == 0) return;
  uint32_t block = 128;
  k_sample_points<<<(n + block - 1) / block, block, 0, stream>>>(nav, o, n, points, flags, skip_if_any,
                                                                 require_all, out_points, out_nodes);
}

}  // namespace lmb
