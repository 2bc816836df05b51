// landmass_b200 — host-side context shared by the C ABI translation units.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "lmb_device.cuh"
#include "lmb_kernels.h"

using namespace lmb;

#define CUDA_TRY(ctx, expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return (ctx)->fail(LMB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  cudaError_t reserve_exact(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  cudaError_t upload(const T* host, size_t n, cudaStream_t s) {
    cudaError_t e = reserve(n ? n : 1);
    if (e != cudaSuccess) return e;
    if (n && host) return cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s);
    return cudaSuccess;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
};


struct lmb_ctx {
  lmb_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;
  bool has_nav = false;
  std::vector<uint32_t> h_island_poly_begin;  // of the current navdata: node id -> island for lmb_navdata_update
  uint32_t sm_count = 148;

  // --- navdata ---
  DevNav nav{};
  DevBuf<DevIsland> d_islands;
  DevBuf<float> d_verts_local, d_verts_world, d_poly_bounds, d_hverts, d_type_cost, d_link_portal,
      d_link_dst_portal, d_link_cost, d_modified_verts;
  DevBuf<uint32_t> d_poly_edge_begin, d_edge_vertex, d_edge_neighbor, d_edge_reverse, d_poly_island,
      d_poly_region, d_poly_type_dense, d_hpoly, d_cell_start, d_cell_polys, d_type_raw, d_link_dst_node,
      d_link_dst_type_dense, d_link_kind, d_link_reverse, d_link_anim_kind, d_link_anim_id,
      d_node_link_begin, d_node_links, d_node_modified, d_modified_edge_begin, d_modified_edges,
      d_modified_vert_begin, d_node_region_number, d_region_root, d_possible_src, d_possible_dst,
      d_possible_kind;
  DevBuf<uint8_t> d_htris, d_type_registered;
  // EdgeRec table followed by the distance table, in ONE allocation: the A* kernels' read-only working set,
  // covered by one L2 access-policy window on the step stream (persisting), so that the per-query arenas
  // streaming through L2 do not evict it
  DevBuf<uint8_t> d_astar_tables;
  DevBuf<float4> d_portal_lr;
  float world_min[2] = {0, 0}, world_max[2] = {0, 0};

  // --- A* ---
  AStarArena arena{};
  DevBuf<uint8_t> d_arena;
  uint32_t kshift = 0, n_ids = 0;              // edge-record id space (see EdgeRec)
  uint32_t arena_node_cap = 0, arena_heap_cap = 0;
  // Small batches on a mesh whose main arena is hashed run on a side arena of dense slots (run_astar)
  AStarArena side_arena{};
  DevBuf<uint8_t> d_side_arena;
  uint32_t side_node_cap = 0, side_heap_cap = 0;
  bool big_heap = false;                       // launch shape for large open lists (k_astar.cu)
  int best_layout = 0;                         // best_estimates layout (k_astar_common.cuh): 0 dense, 1 compact, 2 hashed, 3 none (forest)
  bool nav_is_forest = false;
  DevBuf<AStarCounters> d_counters;
  AStarCounters* h_counters = nullptr;  // pinned

  // --- path pool + persistent per-slot state ---
  DevBuf<uint2> d_pool[2];
  int pool_cur = 0;
  unsigned long long pool_capacity = 0;
  DevBuf<unsigned long long> d_pool_cursor;
  DevBuf<uint32_t> d_slot_path_off, d_slot_path_len, d_slot_cost_hint, d_scan_tmp, d_query_order;
  DevBuf<float> d_desired_move, d_reached_start, d_reached_end;
  DevBuf<uint32_t> d_state, d_reached_link_id;

  // --- per-frame agent inputs (device) ---
  uint32_t n_agents = 0, n_chars = 0;
  DevBuf<uint32_t> d_slot, d_flags, d_reach_condition, d_override_begin, d_override_type,
      d_permitted_begin, d_permitted_kind;
  // per-query filters of lmb_find_paths / lmb_find_straight_paths
  DevBuf<uint32_t> d_qf_override_begin, d_qf_override_type, d_qf_flags, d_qf_permitted_begin, d_qf_permitted_kind;
  DevBuf<float> d_qf_override_cost;
  DevBuf<float> d_position, d_velocity, d_target, d_radius, d_desired_speed, d_max_speed,
      d_reach_distance, d_anim_reach, d_override_cost;
  bool has_reach_condition = false, has_reach_distance = false, has_anim_reach = false,
       has_overrides = false, has_permitted = false;
  DevBuf<float> d_char_position, d_char_velocity, d_char_radius, d_char_point;
  DevBuf<uint32_t> d_char_node;
  // scratch of lmb_sample_points (never the agents' buffers: queries are side-effect free for steps)
  DevBuf<float> d_query_point, d_query_out_point;
  DevBuf<uint32_t> d_query_out_node;
  // sampled
  DevBuf<float> d_agent_point, d_target_point;
  DevBuf<uint32_t> d_agent_node, d_target_node;
  // repath queue + results
  DevBuf<uint32_t> d_queue_count, d_q_agent, d_q_slot, d_q_start_node, d_q_end_node, d_q_status,
      d_q_success, d_q_explored, d_q_path_off, d_q_path_len, d_retry_list;
  DevBuf<float> d_q_start_point, d_q_end_point;
  DevBuf<uint32_t> d_follow_start, d_follow_end, d_presult, d_query_of_agent, d_waypoint_kind, d_waypoint_link;
  DevBuf<uint32_t> d_follow_order, d_follow_hist, d_avoid_flags, d_avoid_pos, d_avoid_order, d_avoid_scan;
  DevBuf<uint8_t> d_follow_key;
  DevBuf<float> d_waypoint_points, d_path_endpoints;
  uint32_t* h_small = nullptr;  // pinned scratch (n_repath etc.)
  // outputs (device)
  DevBuf<float> d_out_velocity, d_out_link_start, d_out_link_end;
  DevBuf<uint32_t> d_out_state, d_out_link_id;
  // avoidance
  DevBuf<AvoidRecord> d_records, d_char_records, d_sorted_records;
  DevBuf<uint32_t> d_cell_count, d_cell_start2, d_avoid_overflow, d_scan_tiles;
  DevBuf<uint8_t> d_avoid_scratch;
  // debug-avoidance export: agents that set LMB_AGENT_KEEP_AVOIDANCE_DATA (in upload order)
  std::vector<uint32_t> h_dbg_row_of;  // [n_agents] row or LMB_NONE; empty = nobody asked
  uint32_t dbg_rows = 0, dbg_lines_per_row = 0;
  DevBuf<uint32_t> d_dbg_row_of, d_dbg_header;
  DevBuf<float4> d_dbg_lines;
  DevBuf<float> d_bounds_tmp;
  uint32_t last_n_results = 0;
  uint32_t avoid_caps[5] = {64, 64, 128, 256, 256};  // neighbours, explore, border edges, cones, lines
  uint64_t avoid_scratch_sig = 0;                    // layout the scratch was last filled for (lmb_avoid_host.cu)

  lmb_step_timings timings{};
  lmb_step_timings query_timings{};  // of the last lmb_find_paths / lmb_find_straight_paths batch
  cudaEvent_t ev[11]{};
  cudaEvent_t ev_q[2]{};
  cudaEvent_t ev_ag[2]{};            // around the library-owned all-gather (lmb_step_sharded)
  uint32_t halo_total = 0, halo_first = 0;
  void* comm = nullptr;  // ncclComm_t of lmb_comm_init (lmb_comm.cu)
  uint32_t comm_ranks = 0, comm_rank = 0;
  bool step_open = false;
  double path_len_estimate = 0.0;
  double explored_estimate = 0.0;  // running mean explored_nodes: the hint of an agent without history  // running mean corridor length, sizes the pool before A*

  int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    error = buf;
    return code;
  }
};

