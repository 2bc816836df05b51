// landmass_b200 — device-side data layout and exact-f32 math.
//
// Every f32 expression on the bit-exact path (sampling, A* costs, funnel sign
// tests) is written with the round-to-nearest intrinsics __fadd_rn/__fmul_rn/
// __fdiv_rn/__fsqrt_rn, which the compiler never contracts into FMA, so results
// match the reference's scalar Rust/glam arithmetic
// independent of -fmad.  The library is additionally built with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/landmass_b200.h"

#define LMB_WARP 32
#define LMB_MAX_TYPES 64u

namespace lmb {

// ---------------------------------------------------------------------------
// exact f32 vector math (glam scalar semantics, left-to-right, no FMA)
// ---------------------------------------------------------------------------
struct V2 {
  float x, y;
};
struct V3 {
  float x, y, z;
};

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ V2 mk2(float x, float y) { return V2{x, y}; }
__device__ __forceinline__ V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return {fadd(a.x, b.x), fadd(a.y, b.y)}; }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return {fsub(a.x, b.x), fsub(a.y, b.y)}; }
__device__ __forceinline__ V2 operator*(V2 a, float s) { return {fmul(a.x, s), fmul(a.y, s)}; }
__device__ __forceinline__ V2 operator/(V2 a, float s) { return {fdiv(a.x, s), fdiv(a.y, s)}; }
__device__ __forceinline__ V2 operator-(V2 a) { return {-a.x, -a.y}; }
__device__ __forceinline__ float dot(V2 a, V2 b) { return fadd(fmul(a.x, b.x), fmul(a.y, b.y)); }
__device__ __forceinline__ float perp_dot(V2 a, V2 b) { return fsub(fmul(a.x, b.y), fmul(a.y, b.x)); }
__device__ __forceinline__ float length_squared(V2 a) { return dot(a, a); }
__device__ __forceinline__ float length(V2 a) { return fsqrt(dot(a, a)); }
__device__ __forceinline__ float distance(V2 a, V2 b) { return length(a - b); }
__device__ __forceinline__ V2 normalize(V2 a) { return a * fdiv(1.0f, length(a)); }

__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)}; }
__device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V2 xy(V3 a) { return {a.x, a.y}; }
__device__ __forceinline__ float dot(V3 a, V3 b) {
  return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ float length_squared(V3 a) { return dot(a, a); }
__device__ __forceinline__ float length(V3 a) { return fsqrt(dot(a, a)); }
__device__ __forceinline__ float distance(V3 a, V3 b) { return length(a - b); }
__device__ __forceinline__ float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {fsub(fmul(a.y, b.z), fmul(b.y, a.z)), fsub(fmul(a.z, b.x), fmul(b.z, a.x)),
          fsub(fmul(a.x, b.y), fmul(b.x, a.y))};
}
__device__ __forceinline__ V3 normalize(V3 a) { return a * fdiv(1.0f, length(a)); }
__device__ __forceinline__ V3 midpoint(V3 a, V3 b) { return (a + b) * 0.5f; }
__device__ __forceinline__ V3 lerp(V3 a, V3 b, float s) { return a * fsub(1.0f, s) + b * s; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}
__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) < __int_as_float(0x7f800000); }
__device__ __forceinline__ V2 normalize_or_zero(V2 a) {
  float rcp = fdiv(1.0f, length(a));
  if (is_finite(rcp) && rcp > 0.0f) return a * rcp;
  return {0.0f, 0.0f};
}
// Quat::from_rotation_z(angle) * v, glam SSE2 evaluation order; (s, c) = sin/cos(angle/2)
// computed once on the host with libm.
__device__ __forceinline__ V3 rotate_z(float s, float c, V3 v) {
  V3 b{0.0f, 0.0f, s};
  float b2 = dot(b, b);
  float k1 = fsub(fmul(c, c), b2);
  float k2 = fmul(dot(v, b), 2.0f);
  float k3 = fmul(c, 2.0f);
  V3 bxv = cross(b, v);
  return (v * k1 + b * k2) + bxv * k3;
}

__device__ __forceinline__ V3 ld3(const float* p, uint32_t i) {
  return {__ldg(p + 3 * (size_t)i), __ldg(p + 3 * (size_t)i + 1), __ldg(p + 3 * (size_t)i + 2)};
}

// ---------------------------------------------------------------------------
// device navigation data
// ---------------------------------------------------------------------------
struct DevIsland {
  float tx, ty, tz;          // translation
  float fs, fc;              // sin/cos(rotation/2)       (apply)
  float is, ic;              // sin/cos(-rotation/2)      (apply_inverse)
  float bmin[3], bmax[3];    // local mesh bounds
  uint32_t bounds_empty;
  uint32_t poly_begin, poly_end;
  uint32_t vert_begin, vert_end;
  uint32_t has_height;
  // sampling grid over polygon AABBs (island-local xy)
  float gx0, gy0, ginv;      // origin, 1/cell
  uint32_t gnx, gny;         // cells
  uint32_t cell_begin;       // offset into cell_start (gnx*gny+1 entries)
};

// One 32-byte record (one sector) per polygon edge = one A* state "in polygon `poly`, standing
// on the midpoint of this edge".  Record ids are the A* state ids: the CSR edge id, or — when
// every polygon has <= 8 edges — `poly << kshift | local edge` (polygons padded to 4 or 8
// records, padding has twin = LMB_NONE), so that a polygon's records are addressable from a
// state id without first reading the state's own record.
struct __align__(32) EdgeRec {
  float mx, my, mz;     // world midpoint: apply(midpoint(local v[e+1], v[e]))
  uint32_t twin;        // record id of the same edge inside the neighbour (= successor state) or LMB_NONE
  uint32_t poly;        // node id owning this edge
  uint32_t first_edge;  // record id of the first edge of `poly`
  uint32_t info;        // bits 0-7 n_edges of poly, 8-15 dense type of poly, 16-23 dense type of
                        // neighbour, bit 24 poly has off-mesh links
  uint32_t aux;         // bits 0-7 index of the twin edge inside the neighbour polygon
};

struct DevNav {
  uint32_t n_islands, n_polys, n_edges, n_vertices, n_links, n_types;
  const DevIsland* islands;
  const float* verts_local;    // [n_vertices*3]
  const float* verts_world;    // [n_vertices*3] apply(local)
  const uint32_t* poly_edge_begin;
  const uint32_t* edge_vertex;
  const uint32_t* edge_neighbor;
  const uint32_t* edge_reverse;
  const float* poly_bounds;    // [n_polys*6]
  const uint32_t* poly_island;
  const uint32_t* poly_region;
  const uint32_t* poly_type_dense;
  const uint32_t* hpoly;
  const float* hverts;
  const uint8_t* htris;
  const uint32_t* cell_start;  // per island blocks
  const uint32_t* cell_polys;
  const EdgeRec* edges;
  const float* edge_dist;      // padded ids only: [n_ids * K] midpoint-to-midpoint distances inside a polygon
  const float4* portal_lr;     // [n_ids * 2] per edge record: world {left.xyz, -}, {right.xyz, -} of the edge as a
                               // portal (left = vertex[e+1], right = vertex[e], nav_mesh.rs:945-950): the funnel walks
                               // a corridor with ONE dependent load per portal instead of four
  uint32_t rec_kshift;         // record id of (node, edge): node << rec_kshift | edge, or (0) poly_edge_begin[node] + edge
  const float* type_cost;      // [n_types] archipelago cost of each dense type (1.0 if unregistered)
  const uint8_t* type_registered;  // [n_types] 1 if present in nav_data.type_index_to_cost
  // links
  const uint32_t* link_dst_node;
  const uint32_t* link_dst_type_dense;
  const float* link_portal;      // [n_links*6] world
  const uint32_t* link_kind;
  const uint32_t* link_reverse;
  const float* link_dst_portal;
  const float* link_cost;
  const uint32_t* link_anim_kind;
  const uint32_t* link_anim_id;
  const uint32_t* node_link_begin;
  const uint32_t* node_links;
  // modified nodes
  const uint32_t* node_modified;        // [n_polys] index into modified arrays or LMB_NONE
  const uint32_t* modified_edge_begin;
  const uint32_t* modified_edges;
  const uint32_t* modified_vert_begin;
  const float* modified_verts;
  // regions
  const uint32_t* node_region_number;
  const uint32_t* region_root;
  uint32_t n_possible_links, n_region_numbers;
  const uint32_t* possible_link_src;
  const uint32_t* possible_link_dst;
  const uint32_t* possible_link_kind;
};

__device__ __forceinline__ V3 island_apply(const DevIsland& is, V3 p) {
  return rotate_z(is.fs, is.fc, p) + V3{is.tx, is.ty, is.tz};
}
__device__ __forceinline__ V3 island_apply_inverse(const DevIsland& is, V3 p) {
  return rotate_z(is.is, is.ic, p - V3{is.tx, is.ty, is.tz});
}

// Sampling-grid cell coordinate; MUST be bit-identical on host (grid build) and device
// (query): monotone in v, so overlapping boxes always share a cell.
__host__ __device__ __forceinline__ int32_t grid_cell(float v, float origin, float inv, uint32_t n) {
#ifdef __CUDA_ARCH__
  float t = __fmul_rn(__fsub_rn(v, origin), inv);
#else
  volatile float d = v - origin;
  volatile float t = d * inv;
#endif
  if (!(t > 0.0f)) return 0;
  if (t >= (float)(n - 1)) return (int32_t)(n - 1);
  return (int32_t)t;
}

}  // namespace lmb
