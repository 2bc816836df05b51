/* TEST INFRASTRUCTURE — C API of the CPU oracle (see oracle/README.md).
 * Single-threaded C++ restatement of the reference's `Archipelago::update`
 * path.  Takes the same flattened inputs as include/landmass_b200.h so the
 * CUDA path and the oracle can be fed identical data. */
#ifndef LANDMASS_ORACLE_H
#define LANDMASS_ORACLE_H
#include "../include/landmass_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(const lmb_navdata_desc* desc, uint32_t flags /* LMB_CONFIG_* */);
void orc_destroy(orc_ctx*);
/* replaces nav data, drops all paths (same contract as lmb_navdata_upload) */
int32_t orc_agent_path_endpoints(orc_ctx*, uint32_t slot, float* out_start_point, float* out_end_point);
int32_t orc_agent_avoidance_constraints(orc_ctx*, uint32_t agent_index, uint32_t capacity, float* out_lines,
                                        uint32_t* out_n_original, uint32_t* out_n_fallback, uint32_t* out_ran);
int32_t orc_agent_waypoints(orc_ctx*, uint32_t capacity, uint32_t* out_kind, float* out_points, uint32_t* out_link,
                            uint32_t* out_count);
int32_t orc_navdata_upload(orc_ctx*, const lmb_navdata_desc* desc);
int32_t orc_navdata_update(orc_ctx*, const lmb_navdata_desc* desc, const lmb_nav_remap* remap);

int32_t orc_sample_points(orc_ctx*, uint32_t n, const float* points, const lmb_options* options,
                          float* out_points, uint32_t* out_nodes);

int32_t orc_find_paths(orc_ctx*, uint32_t n, const uint32_t* start_node, const float* start_point,
                       const uint32_t* end_node, const float* end_point,
                       const uint32_t* override_begin, const uint32_t* override_type,
                       const float* override_cost, uint32_t* out_success, uint32_t* out_explored,
                       uint32_t* out_path_begin, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t path_capacity, const uint32_t* permit_all, const uint32_t* permitted_begin,
                       const uint32_t* permitted_kind);

int32_t orc_find_straight_paths(orc_ctx*, uint32_t n, const uint32_t* start_node, const float* start_point,
                                const uint32_t* end_node, const float* end_point,
                                const uint32_t* override_begin, const uint32_t* override_type,
                                const float* override_cost, uint32_t* out_success,
                                uint32_t* out_step_begin, uint32_t* out_kind, float* out_points,
                                uint32_t* out_link, uint32_t step_capacity, const uint32_t* permit_all,
                                const uint32_t* permitted_begin, const uint32_t* permitted_kind);

int32_t orc_step(orc_ctx*, const lmb_agents_in* agents, const lmb_characters_in* characters,
                 const lmb_options* options, float delta_time, const lmb_agents_out* out);
/* multi-rank form of the step: begin (phases A-C + this rank's records) -> caller all-gathers the
 * halo buffer (HOST memory here) -> end (phase D + outputs).  Same contract as lmb_step_begin/_end. */
int32_t orc_step_begin(orc_ctx*, const lmb_agents_in* agents, const lmb_characters_in* characters,
                       const lmb_options* options, float delta_time, uint32_t n_total, uint32_t first);
int32_t orc_halo_buffer(orc_ctx*, void** host_ptr, uint64_t* record_bytes, uint64_t* n_records);
int32_t orc_step_end(orc_ctx*, const lmb_options* options, float delta_time, const lmb_agents_out* out);
int32_t orc_pathing_results(orc_ctx*, lmb_pathing_result* out, uint32_t capacity,
                            uint32_t* out_count);
int32_t orc_agent_path(orc_ctx*, uint32_t slot, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t capacity, uint32_t* out_len);

/* Funnel (path.rs:227-368) on an explicit flattened path; returns the next
 * waypoint from (start_index,start_point) to (end_index,end_point).
 * out_kind: 0 = Waypoint, 1 = AnimationLink. */
int32_t orc_path_find_index_of_node(orc_ctx*, uint32_t path_len, const uint32_t* nodes, const uint32_t* portals,
                                    uint32_t node, uint32_t reverse, uint32_t* out_found, uint32_t* out_index);
int32_t orc_project_point_to_line_segment(const float* point, const float* start, const float* end,
                                          float* out_point, float* out_fraction);
int32_t orc_next_straight_point(orc_ctx*, uint32_t path_len, const uint32_t* nodes,
                                const uint32_t* portals, uint32_t start_index,
                                const float* start_point, uint32_t end_index, const float* end_point,
                                uint32_t* out_index, uint32_t* out_kind, float* out_point);

/* Generic A* core (astar.rs:126-206) on an adjacency-list problem, for the
 * astar_test.rs goldens.  adjacency of state i = entries adj_begin[i]..adj_begin[i+1]
 * of (adj_cost, adj_action, adj_state).  Returns path length in *out_len (actions
 * in out_actions), *out_found = 0/1, *out_explored. */
int32_t orc_astar_adjacency(uint32_t n_states, const uint32_t* adj_begin, const float* adj_cost,
                            const int32_t* adj_action, const uint32_t* adj_state,
                            const float* heuristic, uint32_t start, uint32_t end,
                            int32_t* out_actions, uint32_t capacity, uint32_t* out_len,
                            uint32_t* out_found, uint32_t* out_explored);

/* Obstacles seen by an agent at `point` on `node` (avoidance.rs:189-471).
 * Output: loops as CSR; out_closed[i] = 1 for Closed, 0 for Open; vertices xy. */
int32_t orc_borders_to_obstacles(orc_ctx*, const float* point, uint32_t node, float distance_limit,
                                 uint32_t* out_n_obstacles, uint32_t* out_vert_begin,
                                 uint32_t* out_closed, float* out_verts, uint32_t obstacle_capacity,
                                 uint32_t vert_capacity);

/* dodgy_2d::Agent::compute_avoiding_velocity restatement (source absent — parity unpinned).
 * agent / neighbours: {px,py,vx,vy,radius,responsibility} (6 floats each).
 * obstacles as CSR of xy vertices with closed flags. */
int32_t orc_compute_avoiding_velocity(const float* agent, uint32_t n_neighbours,
                                      const float* neighbours, uint32_t n_obstacles,
                                      const uint32_t* obstacle_vert_begin,
                                      const uint32_t* obstacle_closed, const float* obstacle_verts,
                                      const float* preferred_velocity, float max_speed,
                                      float time_step, float obstacle_margin, float time_horizon,
                                      float obstacle_time_horizon, float* out_velocity);

#ifdef __cplusplus
}
#endif
#endif
