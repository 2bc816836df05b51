// landmass_b200 — host-callable kernel launchers (one translation unit per kernel family).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lmb_device.cuh"

namespace lmb {

// ---- kernel 1: sampling (k_sample.cu) ----
void launch_sample_points(const DevNav& nav, const lmb_options& o, uint32_t n, const float* points,
                          const uint32_t* flags, uint32_t skip_if_any, uint32_t require_all,
                          float* out_points, uint32_t* out_nodes, cudaStream_t stream);

// ---- kernel 2: batched A* (k_astar.cu) ----
constexpr uint32_t AT_THREADS = 128;  // threads (= concurrent queries) per block, 2 blocks per SM
constexpr uint32_t AT_HEAP_SM = 48;   // unit of the open-list capacity bookkeeping (heap_cap = AT_HEAP_SM + spill growth)
constexpr uint32_t AT_HEAP_SM_MIN = 24;  // smallest number of open-list entries any launch shape keeps in shared
                                         // memory: a slot's spill region holds heap_cap - AT_HEAP_SM_MIN + 2 entries

struct AStarArena {
  uint8_t* base;          // device memory for all slots
  size_t slot_stride;     // bytes per slot
  uint32_t n_slots;       // concurrent queries (= threads launched)
  uint32_t n_ids;         // edge-record ids (CSR: n_edges; padded: n_polys << kshift)
  uint32_t kshift;        // 0 = CSR ids; else a polygon's records are ids [poly << kshift, +1 << kshift)
  uint32_t n_states;      // n_ids + n_links + 2
  uint32_t node_cap;      // nodes per slot
  uint32_t heap_cap;      // open-list entries per query (shared memory + spill)
  size_t nodes_offset;    // byte offset of the node list {state, prev} inside a slot
  size_t heap_offset;     // byte offset of the open-list spill inside a slot
  // best_estimates layout (k_astar_common.cuh): dense f32[n_states] at offset 0, or compact
  uint32_t layout;        // 0 = dense, 1 = compact (per-polygon entries + specials + overflow table),
                          // 2 = hashed (open-addressing table), 3 = none (forest meshes: no region at all)
  uint32_t hash_cap;      // hashed: table entries (power of two, >= 2 x node_cap)
  uint32_t ovf_cap;       // overflow-table entries (power of two)
  size_t special_offset;  // compact: f32[n_links + 2]
  size_t ovf_offset;      // compact: uint2[ovf_cap]
  size_t best_bytes;      // bytes of the best_estimates region (dense array, or polygons + specials)
};

struct AStarQueries {
  uint32_t n;
  const uint32_t* list;        // optional indirection: query ids to run (retry list) or NULL
  const uint32_t* start_node;  // [n]
  const uint32_t* end_node;    // [n]
  const float* start_point;    // [n*3]
  const float* end_point;      // [n*3]
  const uint32_t* owner;       // [n] index into the override / permitted CSR arrays, or NULL (= query id)
  const uint32_t* override_begin;  // CSR by owner, or NULL
  const uint32_t* override_type;   // raw type index
  const float* override_cost;
  const uint32_t* owner_flags;     // agent flags by owner (PERMIT_ALL) or NULL (= all permitted)
  const uint32_t* permitted_begin;
  const uint32_t* permitted_kind;
  const uint32_t* slot;        // [n] agent slot to receive the path, or NULL
  // outputs
  uint32_t* out_status;        // [n] 0 = not run, 1 = done, 2 = retry (arena overflow), 3 = retry (pool
                               //     overflow), 5 = retry with the dense best_estimates layout, 6 = failed, 7 = retry (open-list spill overflow)
  uint32_t* out_success;       // [n]
  uint32_t* out_explored;      // [n]
  uint32_t* out_path_off;      // [n] offset into the path pool
  uint32_t* out_path_len;      // [n]
};

struct PathPool {
  uint2* entries;                   // (node id, portal code)
  unsigned long long* cursor;       // device bump pointer
  unsigned long long capacity;      // entries
  uint32_t* slot_path_off;          // [max_agents]
  uint32_t* slot_path_len;          // [max_agents] 0 = no path
  uint32_t* slot_cost_hint;         // [max_agents] explored_nodes of the slot's previous search (0 = none):
                                    // a scheduling hint only (longest-first), never an input of a result
};

struct AStarCounters {  // device-resident
  unsigned int queue_head;
  unsigned int n_retry_arena;
  unsigned int n_retry_heap;
  unsigned int n_retry_pool;
  unsigned int n_retry_dense;
  unsigned int n_failed;
  unsigned int n_done;
  unsigned long long explored_total;
  unsigned long long path_nodes_total;
  unsigned long long heap_max_total;  // sum over finished queries of their largest open list
};

// flags: lmb_config.flags (LMB_CONFIG_ASTAR_* launch-shape test hooks).  Returns the kernel family launched:
// 1 = thread-per-query (k_astar.cu), 2 = warp-per-query (k_astar_warp.cu), 0 = nothing.
uint32_t launch_astar(const DevNav& nav, const AStarArena& arena, const AStarQueries& q, const PathPool& pool,
                      AStarCounters* counters, const uint32_t* type_raw, bool big_heap, uint32_t flags,
                      uint32_t shape, cudaStream_t stream);
// order[] = query indices 0..n-1 sorted by descending cost hint of their agent slot (bucketed)
void launch_order_queries(uint32_t n, const uint32_t* q_slot, const uint32_t* slot_cost_hint, uint32_t unknown_hint,
                          uint32_t* order, cudaStream_t stream);
void launch_build_edge_dist(const EdgeRec* recs, uint32_t n_ids, uint32_t kshift, float* out, cudaStream_t stream);
uint32_t astar_resident_queries_per_sm(bool big_heap, uint32_t shape);
void launch_arena_init(const AStarArena& arena, cudaStream_t stream);

// ---- kernel 3: repath decision + funnel / follow (k_follow.cu) ----
struct AgentArrays {
  uint32_t n;
  const uint32_t* slot;
  const uint32_t* flags;
  const float* position;
  const float* velocity;
  const float* target;
  const float* radius;
  const float* desired_speed;
  const float* max_speed;
  const uint32_t* reach_condition;
  const float* reach_distance;
  const float* anim_link_reach_distance;
  // sampled (phase A outputs)
  const float* agent_point;
  const uint32_t* agent_node;
  const float* target_point;
  const uint32_t* target_node;
};

struct AgentPersist {  // indexed by slot
  float* desired_move;       // [max_agents*3]
  uint32_t* state;           // [max_agents]
  uint32_t* reached_link_id;
  float* reached_link_start;
  float* reached_link_end;
  float* path_endpoints;     // [max_agents*6] Path.start_point, Path.end_point (path.rs:24-27) of the stored path
};

struct RepathQueue {
  uint32_t* count;       // device counter
  uint32_t* agent;       // [n] dense agent index of each query
  uint32_t* start_node;
  uint32_t* end_node;
  float* start_point;
  float* end_point;
  uint32_t* slot;
};

// per-agent frame scratch
struct FollowScratch {
  uint32_t* follow_start;  // [n] flat path index of the agent node, LMB_NONE if no follow indices
  uint32_t* follow_end;    // [n]
  uint32_t* presult;       // [n] bit31 repathed this frame, bit30 success, low 30 bits explored
  uint32_t* query_of_agent;  // [n] query id or LMB_NONE
  // the frame's next step of the straight path per agent (what debug drawing shows, debug.rs:386-412):
  // kind LMB_NONE = none this frame, 0 = Waypoint, 1 = AnimationLink
  uint32_t* waypoint_kind;   // [n]
  uint32_t* waypoint_link;   // [n] animation link id of an AnimationLink step
  float* waypoint_points;    // [n*6] point (or link start point) xyz, link end point xyz
  // k_follow's processing order (longest remaining corridor first, k_follow_order_*), or nullptr = by index
  const uint32_t* order;     // [n]
};
constexpr uint32_t FOLLOW_BUCKETS = 64;

void launch_reset_slots(const AgentArrays& ag, const AgentPersist& persist, const PathPool& pool,
                        cudaStream_t stream);
void launch_repath_decide(const DevNav& nav, const AgentArrays& ag, const AgentPersist& persist,
                          const PathPool& pool, const RepathQueue& queue, const FollowScratch& scratch,
                          cudaStream_t stream);
void launch_follow(const DevNav& nav, const AgentArrays& ag, const AgentPersist& persist,
                   const PathPool& pool, const AStarQueries& q, const FollowScratch& scratch,
                   cudaStream_t stream);
// order[] = the agents sorted by the length of corridor the funnel may scan, longest first (counting sort over
// half-octave buckets); key8 [n] bytes, hist [2 * FOLLOW_BUCKETS] words.  Three launches.
void launch_follow_order(const AgentArrays& ag, const PathPool& pool, const AStarQueries& q, const FollowScratch& scratch,
                         uint8_t* key8, uint32_t* hist, uint32_t* order, cudaStream_t stream);
void launch_gather_outputs(const AgentArrays& ag, const AgentPersist& persist, float* out_velocity,
                           uint32_t* out_state, uint32_t* out_link_id, float* out_link_start,
                           float* out_link_end, cudaStream_t stream);
void launch_straight_paths(const DevNav& nav, const PathPool& pool, const AStarQueries& q,
                           const uint32_t* step_begin, uint32_t* step_count, uint32_t* out_kind,
                           float* out_points, uint32_t* out_link, cudaStream_t stream);
void launch_drop_all_paths(const PathPool& pool, uint32_t max_agents, cudaStream_t stream);
void launch_remap_paths(const PathPool& pool, uint32_t max_agents, const uint32_t* old_begin, uint32_t n_old_islands,
                        const uint32_t* new_begin_of_old, const uint32_t* link_map, uint32_t n_old_links,
                        cudaStream_t stream);

// path pool compaction
void launch_pool_scan(const PathPool& src, uint32_t max_agents, uint32_t* scan_tmp, uint32_t* tile_tmp,
                      cudaStream_t stream);
void launch_pool_copy(const PathPool& src, uint2* dst_entries, uint32_t max_agents, const uint32_t* scan_tmp,
                      cudaStream_t stream);

// ---- kernel 4: avoidance (k_avoid.cu) ----
struct AvoidRecord {  // 32 B per sampled agent: the per-frame halo exchanged between GPUs
  float px, py, pz;
  float vx, vy;
  float radius;
  float responsibility;
  uint32_t valid;
};

struct AvoidGrid {
  float ox, oy, inv_cell;
  uint32_t nx, ny;
  uint32_t* cell_count;   // [nx*ny + 1]
  uint32_t* cell_start;   // [nx*ny + 1]
  AvoidRecord* sorted_rec;  // [n_records] the records themselves in cell order, `valid` replaced by the
                            // record's index (the cells of a grid row are contiguous: an agent's candidates of one
                            // row are then one contiguous run of 32-byte records)
  uint32_t* scan_tmp;     // [exclusive_scan_tmp_words(nx*ny + 1)]
  uint32_t capacity_cells;
};

// The order in which k_avoidance takes a rank's agents: by index (mode 0), or in the grid's cell order, so that the
// threads of a warp stand next to each other (the same candidate rows, about as many neighbours).  Mode 1: every
// record is the rank's own and the cell-ordered record copies ARE the order; mode 2 (agents sharded over ranks):
// the rank's own records picked out of the cell order (flags, scan, scatter).
struct AvoidOrder {
  uint32_t mode;
  uint32_t* flags;   // [n_records + 1]
  uint32_t* pos;     // [n_records + 1] exclusive scan of flags; pos[n_records] = how many
  uint32_t* order;   // [n_agents] local dense agent indices
  uint32_t* scan_tmp;
};

struct AvoidScratch {
  // per-agent variable-length scratch, sized by `per_agent_*`
  uint8_t* base;
  size_t per_agent_bytes;
  uint32_t agent_stride;  // agents per chunk rounded up to 32: the arrays are interleaved by agent (k_avoid.cu)
  uint32_t max_neighbours;
  uint32_t max_lines;
  uint32_t max_cones;
  uint32_t max_border_edges;
  uint32_t max_explore;
  uint32_t* overflow;  // device flag
};

// debug-avoidance export (debug.rs:415-503): the ORCA constraint lines of the agents that set
// LMB_AGENT_KEEP_AVOIDANCE_DATA, one row per such agent.  row_of == nullptr: nobody asked.
struct AvoidDebug {
  const uint32_t* row_of;  // [n_agents] row of the agent or LMB_NONE
  uint32_t* header;        // [rows][4] {ran this step, n original, n fallback, fell back}
  float4* lines;           // [rows][lines_per_row] {point.xy, direction.xy}: originals, then fallbacks
  uint32_t lines_per_row;
};

void launch_build_records(const AgentArrays& ag, const AgentPersist& persist, const lmb_options& o,
                          AvoidRecord* records, cudaStream_t stream);
void launch_character_records(uint32_t n, const float* velocity, const float* radius, const float* points,
                              const uint32_t* nodes, AvoidRecord* records, cudaStream_t stream);
void launch_avoidance(const DevNav& nav, const AgentArrays& ag, const AgentPersist& persist,
                      const lmb_options& o, float delta_time, uint32_t first_agent_global,
                      uint32_t n_records, const AvoidRecord* records, uint32_t n_characters,
                      const AvoidRecord* char_records, AvoidGrid& grid, const AvoidScratch& scratch,
                      float* bounds_tmp, cudaStream_t stream);

// exclusive scan (u32), in place allowed; tmp = exclusive_scan_tmp_words(n) words (NULL: single block)
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t stream);
uint32_t exclusive_scan_tmp_words(uint32_t n);
void launch_integrate(uint32_t n, float dt, const float* out_velocity, float* position, float* velocity,
                      uint32_t* flags, cudaStream_t stream);

}  // namespace lmb
