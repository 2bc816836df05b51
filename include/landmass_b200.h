/*
 * landmass_b200.h — C ABI of the B200-native landmass agent step.
 *
 * This is the drop-in boundary for ONE hot path of the landmass crate:
 * `Archipelago::update` (reference: crates/landmass/src/lib.rs:270-534).
 * The reference has no FFI of its own (100 % safe Rust); the entry points below
 * are what a thin `landmass-b200-sys` crate would bind so that the public Rust
 * API (`Archipelago`, `Agent`, `Character`, `Island`, `ValidNavigationMesh`,
 * `ArchipelagoOptions`, `AgentState`) stays unchanged.  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int32 status: 0 = ok, <0 = error (lmb_status);
 *     no exception / abort crosses the ABI; `lmb_last_error` gives the text.
 *   - all pointers are caller-owned HOST memory, only read/written during the
 *     call; the library copies what it keeps.  All output buffers are
 *     caller-allocated with explicit capacities.
 *   - all coordinates are landmass "standard" coordinates (after
 *     `CoordinateSystem::to_landmass`, reference coords.rs:6-24): x right,
 *     y forward, z up; f32.
 *   - "node id" = global polygon index = island_poly_begin[island] + polygon.
 *   - LMB_NONE (0xFFFFFFFF) is the "no value" sentinel for u32 fields.
 */
#ifndef LANDMASS_B200_H
#define LANDMASS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMB_ABI_VERSION 4u
#define LMB_NONE 0xFFFFFFFFu

typedef enum lmb_status {
  LMB_OK = 0,
  LMB_ERR_INVALID_ARGUMENT = -1,
  LMB_ERR_CUDA = -2,
  LMB_ERR_NO_NAVDATA = -3,
  LMB_ERR_CAPACITY = -4,
  LMB_ERR_UNSUPPORTED = -5,
  LMB_ERR_INTERNAL = -6
} lmb_status;

/* AgentState codes — reference agent.rs:24-47 (same order as the Rust enum). */
typedef enum lmb_agent_state {
  LMB_STATE_IDLE = 0,
  LMB_STATE_REACHED_TARGET = 1,
  LMB_STATE_REACHED_ANIMATION_LINK = 2,
  LMB_STATE_USING_ANIMATION_LINK = 3,
  LMB_STATE_MOVING = 4,
  LMB_STATE_AGENT_NOT_ON_NAV_MESH = 5,
  LMB_STATE_TARGET_NOT_ON_NAV_MESH = 6,
  LMB_STATE_NO_PATH = 7,
  LMB_STATE_PAUSED = 8
} lmb_agent_state;

/* TargetReachedCondition tags — reference agent.rs:136-160. */
#define LMB_REACH_DISTANCE 0u
#define LMB_REACH_VISIBLE_AT_DISTANCE 1u
#define LMB_REACH_STRAIGHT_PATH_DISTANCE 2u

/* Per-agent flag bits (lmb_agents_in.flags). */
#define LMB_AGENT_HAS_TARGET 1u            /* agent.current_target.is_some() */
#define LMB_AGENT_PAUSED 2u                /* agent.paused */
#define LMB_AGENT_USING_ANIMATION_LINK 4u  /* agent.using_animation_link */
#define LMB_AGENT_RESET 8u                 /* slot was (re)created: drop the device-side path/state */
#define LMB_AGENT_PERMIT_ALL_LINKS 16u     /* PermittedAnimationLinks::All */
#define LMB_AGENT_KEEP_AVOIDANCE_DATA 32u  /* agent.keep_avoidance_data (agent.rs:87-90, feature debug-avoidance) */

/* Off-mesh link kinds — reference nav_data.rs:96-117. */
#define LMB_LINK_BOUNDARY 0u
#define LMB_LINK_ANIMATION 1u

/* Portal code in a path entry: edge index of the node, or LMB_PORTAL_LINK|link
 * index for the off-mesh link that leaves the node, or LMB_NONE for the last
 * node of the path.  (reference path.rs:14-55: IslandSegment.portal_edge_index
 * / OffMeshLinkSegment flattened into one list in PathIndex order.) */
#define LMB_PORTAL_LINK 0x80000000u

/*
 * ArchipelagoOptions (reference lib.rs:63-79) with the point sample distance
 * already evaluated to CorePointSampleDistance (reference coords.rs:209-226).
 */
typedef struct lmb_options {
  float horizontal_distance;
  float distance_above;
  float distance_below;
  float vertical_preference_ratio;
  float neighbourhood;
  float avoidance_time_horizon;
  float obstacle_avoidance_time_horizon;
  float reached_destination_avoidance_responsibility;
} lmb_options;

/*
 * Flattened, validated navigation data: what `NavigationData` holds after
 * `nav_data.update()` (reference nav_data.rs:27-62), i.e. the islands'
 * `ValidNavigationMesh`es (nav_mesh.rs:483-583) concatenated in
 * `DenseSlotMap` iteration order, plus off-mesh links, modified nodes and
 * region connectivity.  Every array is SoA, indices are u32.
 */
typedef struct lmb_navdata_desc {
  uint32_t struct_size; /* = sizeof(lmb_navdata_desc), ABI guard */

  /* ---- islands, in DenseSlotMap iteration order (tie-break order of
   *      nav_data.rs:275,309-313) ---- */
  uint32_t n_islands;
  const float* island_translation;    /* [n_islands*3] Transform.translation (to_landmass) */
  const float* island_rotation;       /* [n_islands]   Transform.rotation (radians, about z) */
  const float* island_bounds;         /* [n_islands*6] local mesh_bounds min xyz,max xyz; min.x>max.x = Empty */
  const uint32_t* island_poly_begin;  /* [n_islands+1] first node id of each island */
  const uint32_t* island_vert_begin;  /* [n_islands+1] first vertex of each island */

  /* ---- vertices (island-local coordinates) ---- */
  uint32_t n_vertices;
  const float* vertices; /* [n_vertices*3] */

  /* ---- polygons (nodes); edge e of a polygon runs vertex[e] -> vertex[e+1] ---- */
  uint32_t n_polys;
  const uint32_t* poly_edge_begin; /* [n_polys+1] CSR into the edge arrays */
  uint32_t n_edges;
  const uint32_t* edge_vertex;   /* [n_edges] GLOBAL vertex index of the edge's first vertex */
  const uint32_t* edge_neighbor; /* [n_edges] node id across the edge, or LMB_NONE (Connectivity.polygon_index) */
  const uint32_t* edge_reverse;  /* [n_edges] Connectivity.reverse_edge (edge index inside the neighbour) */
  const uint32_t* poly_type;     /* [n_polys] ValidPolygon.type_index */
  const uint32_t* poly_region;   /* [n_polys] ValidPolygon.region (per island) */
  const float* poly_bounds;      /* [n_polys*6] ValidPolygon.bounds (island-local) */
  const float* poly_center;      /* [n_polys*3] ValidPolygon.center (island-local) */

  /* ---- optional height meshes (nav_mesh.rs:503-521) ---- */
  const uint8_t* island_has_height; /* [n_islands] or NULL */
  const uint32_t* hpoly;            /* [n_polys*4] base_vertex (GLOBAL into hverts), vertex_count,
                                       base_triangle (GLOBAL into htris), triangle_count; NULL if none */
  uint32_t n_hverts;
  const float* hverts; /* [n_hverts*3] island-local */
  uint32_t n_htris;
  const uint8_t* htris; /* [n_htris*3] indices relative to base_vertex */

  /* ---- type index -> cost (nav_data.rs:34); unlisted types cost 1.0 ---- */
  uint32_t n_type_costs;
  const uint32_t* type_cost_index;
  const float* type_cost_value;

  /* ---- off-mesh links (nav_data.rs:78-117), canonical ascending order ---- */
  uint32_t n_links;
  const uint32_t* link_src_node;   /* [n_links] node the link leaves from */
  const uint32_t* link_dst_node;   /* [n_links] OffMeshLink.destination_node */
  const uint32_t* link_dst_type;   /* [n_links] OffMeshLink.destination_type_index */
  const float* link_portal;        /* [n_links*6] world-space portal .0 xyz, .1 xyz */
  const uint32_t* link_kind;       /* [n_links] LMB_LINK_* */
  const uint32_t* link_reverse;    /* [n_links] BoundaryLink.reverse_link (link index) or LMB_NONE */
  const float* link_dst_portal;    /* [n_links*6] AnimationLink.destination_portal (unused for boundary) */
  const float* link_cost;          /* [n_links] AnimationLink.cost */
  const uint32_t* link_anim_kind;  /* [n_links] AnimationLink.kind */
  const uint32_t* link_anim_id;    /* [n_links] caller's id of the AnimationLink */
  const uint32_t* node_link_begin; /* [n_polys+1] CSR: links leaving each node ... */
  const uint32_t* node_links;      /* [n_links]   ... ascending link index */

  /* ---- modified nodes (nav_data.rs:119-132), CSR ---- */
  uint32_t n_modified;
  const uint32_t* modified_node;        /* [n_modified] node id, ascending */
  const uint32_t* modified_edge_begin;  /* [n_modified+1] */
  const uint32_t* modified_edges;       /* [2*n] (left,right); < island vertex count: island-local vertex
                                           index, else new vertex (index - vertex count) */
  const uint32_t* modified_vert_begin;  /* [n_modified+1] */
  const float* modified_verts;          /* [2*n] ModifiedNode.new_vertices (world xy) */

  /* ---- region connectivity (nav_data.rs:38-48,1022-1087) ---- */
  const uint32_t* node_region_number; /* [n_polys] region_id_to_number[(island, region)] or LMB_NONE */
  uint32_t n_region_numbers;
  const uint32_t* region_root;        /* [n_region_numbers] representative of region_connections */
  uint32_t n_possible_links;          /* region_number_to_possible_links, grouped by source, in push order */
  const uint32_t* possible_link_src;  /* region number */
  const uint32_t* possible_link_dst;  /* region number (PossibleRegionLink.destination_region) */
  const uint32_t* possible_link_kind; /* PossibleRegionLink.link_kind */
} lmb_navdata_desc;

/*
 * Per-frame agent inputs, SoA, in `DenseSlotMap` iteration order
 * (reference lib.rs:57; Agent pub fields agent.rs:50-92).
 */
typedef struct lmb_agents_in {
  uint32_t n;
  const uint32_t* slot;     /* [n] stable slot id < max_agents; keys the device-side path state */
  const uint32_t* flags;    /* [n] LMB_AGENT_* */
  const float* position;    /* [n*3] to_landmass(agent.position) */
  const float* velocity;    /* [n*3] to_landmass(agent.velocity) */
  const float* target;      /* [n*3] to_landmass(current_target) (ignored without HAS_TARGET) */
  const float* radius;      /* [n] */
  const float* desired_speed; /* [n] */
  const float* max_speed;   /* [n] */
  const uint32_t* reach_condition; /* [n] LMB_REACH_* ; NULL = all Distance */
  const float* reach_distance;     /* [n] < 0 or NaN = None (use radius); NULL = all None */
  const float* anim_link_reach_distance; /* [n] <0 = None; NULL = all None (agent.rs:312-323) */
  /* per-agent override_type_index_to_cost (agent.rs:92) as CSR; NULL = none */
  const uint32_t* override_begin; /* [n+1] */
  const uint32_t* override_type;
  const float* override_cost;
  /* PermittedAnimationLinks::Kinds (agent.rs:170-187) as CSR; used only when
   * LMB_AGENT_PERMIT_ALL_LINKS is clear; NULL = none permitted */
  const uint32_t* permitted_begin; /* [n+1] */
  const uint32_t* permitted_kind;
} lmb_agents_in;

/* Characters — reference character.rs:14-21. */
typedef struct lmb_characters_in {
  uint32_t n;
  const float* position; /* [n*3] */
  const float* velocity; /* [n*3] */
  const float* radius;   /* [n] */
} lmb_characters_in;

/* Per-frame outputs, same order as lmb_agents_in. */
typedef struct lmb_agents_out {
  float* desired_velocity; /* [n*3] agent.current_desired_move in landmass coords (z = 0) */
  uint32_t* state;         /* [n] lmb_agent_state */
  /* ReachedAnimationLink (agent.rs:112-119); link id LMB_NONE = None. May be NULL. */
  uint32_t* reached_link_id; /* [n] */
  float* reached_link_start; /* [n*3] */
  float* reached_link_end;   /* [n*3] */
} lmb_agents_out;

/* PathingResult — reference lib.rs:538-547; `agent` is the index in lmb_agents_in order. */
typedef struct lmb_pathing_result {
  uint32_t agent;
  uint32_t success;
  uint32_t explored_nodes;
} lmb_pathing_result;

#define LMB_CONFIG_NO_AVOIDANCE 1u /* skip phase D (testing phases A-C in isolation) */
#define LMB_CONFIG_DEBUG_LOG 2u    /* one stderr line per A* arena (re)allocation: layout, slots, sizes */
/* Test hooks of the A* kernels' data layouts and launch shapes (tests/test_astar_layouts.py).  None of
 * them changes a result; production callers leave them clear and the library chooses. */
#define LMB_CONFIG_ASTAR_COMPACT 4u        /* start with the per-polygon best_estimates layout even off a forest */
#define LMB_CONFIG_ASTAR_DENSE 8u          /* dense best_estimates whatever the mesh */
#define LMB_CONFIG_ASTAR_HASHED 16u        /* hashed best_estimates with a 2048-node arena */
#define LMB_CONFIG_ASTAR_LARGE_SHAPES 32u  /* thread-per-query large-batch launch shapes for every batch */
#define LMB_CONFIG_ASTAR_THREAD_ONLY 64u   /* never launch the warp-per-query (low latency) search kernel */
#define LMB_CONFIG_ASTAR_WARP_ONLY 128u    /* launch only the warp-per-query search kernel */
#define LMB_CONFIG_ASTAR_NO_FOREST 256u    /* do not use the best_estimates-free layout of forest meshes */
#define LMB_CONFIG_NO_L2_WINDOW 512u       /* do not pin the A* tables in L2 (access-policy window; A/B hook) */
#define LMB_CONFIG_ASTAR_SMALL_HEAP 1024u  /* start with a 64-entry open list per query: forces the in-place spill,
                                              its overflow (status 7) and the doubling retry at test sizes */
#define LMB_CONFIG_ASTAR_NO_PAIR 2048u     /* warp-per-query launches never use a second (heap) warp per query */
#define LMB_CONFIG_ORDER_ALWAYS 4096u      /* k_follow / k_avoidance take the agents in their work order (longest corridor
                                              first / grid cell order) at every size, not only from 8192 agents on */

typedef struct lmb_config {
  uint32_t struct_size;
  int32_t device;            /* CUDA device ordinal */
  uint32_t max_agents;       /* capacity of slot ids */
  uint32_t astar_slots;      /* A* arena slots = queries in flight; 0 = auto (up to 256 per SM) */
  uint64_t scratch_bytes;    /* byte budget of the A* arena; 0 = auto (40 % of free HBM, <= 64 GB) */
  uint32_t flags;            /* LMB_CONFIG_* */
  uint32_t astar_shape;      /* tuning hook: launch shape of large A* batches on forest meshes
                                (0 = library default; k_astar.cu launch_astar lists the others) */
} lmb_config;

/* Device timings (milliseconds, CUDA events on the step stream) of the last step. */
typedef struct lmb_step_timings {
  float h2d_ms;
  float sample_ms;
  float repath_decide_ms;
  float astar_ms;
  float follow_ms;
  float avoidance_ms;
  float d2h_ms;
  float total_ms;
  uint32_t n_repath;          /* A* queries this step */
  uint32_t kernel_launches;   /* kernels launched by the step */
  uint64_t explored_nodes;    /* sum over this step's queries */
  uint64_t path_nodes;        /* sum of corridor lengths produced this step */
  float allgather_ms;         /* lmb_step_sharded: device time of the ncclAllGather (0 otherwise) */
  uint32_t astar_launch_kind; /* A* launches of the step, or-ed: 1 thread-per-query, 2 warp-per-query,
                                 4 warp pair per query (search warp + heap warp) */
} lmb_step_timings;

typedef struct lmb_ctx lmb_ctx;

uint32_t lmb_abi_version(void);
int32_t lmb_create(const lmb_config* config, lmb_ctx** out_ctx);
void lmb_destroy(lmb_ctx* ctx);
/* Text of the last error on this context (or of a failed lmb_create when ctx == NULL). */
const char* lmb_last_error(const lmb_ctx* ctx);

/* Replaces the navigation data.  Call when `nav_data.update()` (nav_data.rs:1158-1180)
 * reports a change.  All device-side paths are dropped (agents with a target repath) —
 * i.e. every island counts as invalidated; use lmb_navdata_update to keep the others. */
int32_t lmb_navdata_upload(lmb_ctx* ctx, const lmb_navdata_desc* desc);

/* What `NavigationData::update` returns (nav_data.rs:326-350: dropped off-mesh links, changed
 * islands), expressed against the PREVIOUS upload's flattening so that stored paths can be
 * carried over: node ids and link indices change whenever an island or link is added or removed. */
typedef struct lmb_nav_remap {
  uint32_t struct_size;
  uint32_t n_old_islands;     /* n_islands of the previous upload */
  const uint32_t* island_map; /* [n_old_islands] index of the same, UNCHANGED island in `desc`, or
                                 LMB_NONE when it was removed or dirty (invalidated_islands) */
  uint32_t n_old_links;       /* n_links of the previous upload */
  const uint32_t* link_map;   /* [n_old_links] index of the same off-mesh link in `desc`, or LMB_NONE
                                 when it was dropped (invalidated_off_mesh_links) */
} lmb_nav_remap;

/* lmb_navdata_upload that keeps the paths `Path::is_valid` (path.rs:381-400) keeps: a stored path
 * survives when none of its island segments lies on an invalidated island and none of its off-mesh
 * link segments uses a dropped link; its node ids / link indices are rewritten to the new
 * flattening.  Every other path is marked invalid and handled by the next step exactly like the
 * reference (lib.rs:350-358, agent.rs:454-456): paused agents lose it, agents with a target
 * re-plan.  remap == NULL is lmb_navdata_upload. */
int32_t lmb_navdata_update(lmb_ctx* ctx, const lmb_navdata_desc* desc, const lmb_nav_remap* remap);

/* One `Archipelago::update(delta_time)` (lib.rs:270-534) with HOST buffers:
 * H2D of inputs, phases A-D on the device, D2H of outputs; synchronous. */
int32_t lmb_step(lmb_ctx* ctx, const lmb_agents_in* agents, const lmb_characters_in* characters,
                 const lmb_options* options, float delta_time, const lmb_agents_out* out);

/* The same step split so inputs can stay resident in HBM:
 * upload -> (step_resident)* -> download. */
int32_t lmb_agents_upload(lmb_ctx* ctx, const lmb_agents_in* agents,
                          const lmb_characters_in* characters);
int32_t lmb_step_resident(lmb_ctx* ctx, const lmb_options* options, float delta_time);
int32_t lmb_agents_download(lmb_ctx* ctx, const lmb_agents_out* out);
/* For a device-resident loop: what the caller's own movement system does between two updates (README example
 * README.md:55-131 — velocity = the desired velocity just computed, position += velocity * delta_time) applied
 * to the resident inputs, and LMB_AGENT_RESET cleared.  Enqueued on the library stream; no synchronisation. */
int32_t lmb_agents_integrate(lmb_ctx* ctx, float delta_time);
/* The resident positions (as uploaded, or as advanced by lmb_agents_integrate): [n*3] floats. */
int32_t lmb_agents_positions(lmb_ctx* ctx, float* out_position, uint32_t capacity);

/* Multi-GPU form of the step: agents are sharded over ranks by contiguous
 * ranges of the global agent order, navigation data is replicated, and the only exchange is
 * the per-agent avoidance record (32 bytes) between phase C and phase D.
 *   lmb_step_begin : phases A-C for this rank's agents (those given to lmb_agents_upload),
 *                    then writes their records at [first_agent, first_agent + n) of the halo
 *                    buffer; returns after the stream is idle.
 *   lmb_halo_buffer: device pointer / layout of the halo buffer (n_total records); the caller
 *                    all-gathers it in place (e.g. ncclAllGather) and synchronises.
 *   lmb_step_end   : phase D for this rank's agents against all n_total records, outputs.
 * lmb_step_resident(ctx, o, dt) == begin(n, 0) + end. */
int32_t lmb_step_begin(lmb_ctx* ctx, const lmb_options* options, float delta_time,
                       uint32_t n_total_agents, uint32_t first_agent);
int32_t lmb_halo_buffer(lmb_ctx* ctx, void** device_ptr, uint64_t* record_bytes, uint64_t* n_records);
int32_t lmb_step_end(lmb_ctx* ctx, const lmb_options* options, float delta_time);

/* The same exchange done by the library, for callers without a collective of their
 * own: the library opens libnccl.so.2 at run time and owns one communicator per context.
 *   lmb_comm_unique_id : rank 0 makes the 128-byte ncclUniqueId; the caller ships it to every rank
 *                        by whatever channel it has (the bench uses torch.distributed broadcast).
 *   lmb_comm_init      : ncclCommInitRank on the context's device (collective over all ranks).
 *   lmb_step_sharded   : phases A-C, ncclAllGather of the records on the library's stream, phase D,
 *                        outputs — no host synchronisation between the phases.  Every rank passes
 *                        the same `agents_per_rank` >= its own agent count; rank r's records live at
 *                        [r * agents_per_rank, +n_r) of the halo buffer and the rest of its share is
 *                        marked invalid, so ranks need not hold equal numbers of agents.
 * lmb_get_step_timings().allgather_ms is the device time of the collective. */
int32_t lmb_comm_unique_id(lmb_ctx* ctx, uint8_t out_id[128]);
int32_t lmb_comm_init(lmb_ctx* ctx, const uint8_t id[128], uint32_t n_ranks, uint32_t rank);
int32_t lmb_step_sharded(lmb_ctx* ctx, const lmb_options* options, float delta_time, uint32_t agents_per_rank);

/* `Archipelago::get_pathing_results` (lib.rs:231-233): results of the last step in agent order. */
int32_t lmb_pathing_results(lmb_ctx* ctx, lmb_pathing_result* out, uint32_t capacity,
                            uint32_t* out_count);

/* Reads back `agent.current_path` of one slot (debug.rs / tests): flattened
 * (node id, portal code) pairs.  *out_len = 0 when the agent has no path;
 * LMB_ERR_CAPACITY (with *out_len set) if capacity is too small. */
int32_t lmb_agent_path(lmb_ctx* ctx, uint32_t slot, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t capacity, uint32_t* out_len);

/* Batched `NavigationData::sample_point` (nav_data.rs:269-324).
 * out_nodes[i] = LMB_NONE when nothing is in range. */
int32_t lmb_sample_points(lmb_ctx* ctx, uint32_t n, const float* points, const lmb_options* options,
                          float* out_points, uint32_t* out_nodes);

/* Batched `pathfinding::find_path` (pathfinding.rs:277-382).  Query i goes from
 * (start_node[i], start_point[i]) to (end_node[i], end_point[i]).  Overrides are CSR per query
 * (NULL = none).  permitted_animation_links per query: permit_all[i] != 0 is
 * PermittedAnimationLinks::All, otherwise Kinds = row i of the permitted CSR; permit_all == NULL
 * means All for every query (the CSR is then ignored and may be NULL).
 * Output corridor i = out_nodes/out_portals[out_path_begin[i] .. out_path_begin[i+1]);
 * out_success[i] = 0 leaves an empty corridor.  path_capacity is the size of
 * out_nodes/out_portals; LMB_ERR_CAPACITY if the paths do not fit. */
int32_t lmb_find_paths(lmb_ctx* ctx, uint32_t n, const uint32_t* start_node, const float* start_point,
                       const uint32_t* end_node, const float* end_point,
                       const uint32_t* override_begin, const uint32_t* override_type,
                       const float* override_cost, uint32_t* out_success, uint32_t* out_explored,
                       uint32_t* out_path_begin, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t path_capacity, const uint32_t* permit_all, const uint32_t* permitted_begin,
                       const uint32_t* permitted_kind);

/* Batched `Archipelago::find_path` of the query API (query.rs:185-273): A* followed by the
 * full straight path (repeated funnel).  Steps of query i are
 * [out_step_begin[i], out_step_begin[i+1]); the first step is the start point.  Per step:
 * out_kind = 0 Waypoint / 1 AnimationLink, out_points = 6 floats (waypoint xyz twice, or link
 * start xyz + link end xyz), out_link = animation link id or LMB_NONE.  A failed query
 * (FindPathError::NoPathFound) has out_success = 0 and no steps. */
int32_t lmb_find_straight_paths(lmb_ctx* ctx, uint32_t n, const uint32_t* start_node,
                                const float* start_point, const uint32_t* end_node,
                                const float* end_point, const uint32_t* override_begin,
                                const uint32_t* override_type, const float* override_cost,
                                uint32_t* out_success, uint32_t* out_step_begin, uint32_t* out_kind,
                                float* out_points, uint32_t* out_link, uint32_t step_capacity,
                                const uint32_t* permit_all, const uint32_t* permitted_begin,
                                const uint32_t* permitted_kind);

/* Read-backs for `debug::draw_archipelago_debug` (debug.rs:89-413), both cold:
 * the start and end point of the path stored for `slot` (Path.start_point / end_point,
 * path.rs:24-27; meaningful while lmb_agent_path reports a path), and, for every agent of the last
 * step in its order, the next step of the straight path that step computed (debug.rs:386-412):
 * out_kind LMB_NONE = none (no path, paused, using a link), 0 = Waypoint, 1 = AnimationLink;
 * out_points = 6 floats (waypoint xyz twice, or link start xyz + link end xyz); out_link = the
 * animation link id of an AnimationLink step. */
int32_t lmb_agent_path_endpoints(lmb_ctx* ctx, uint32_t slot, float* out_start_point, float* out_end_point);
int32_t lmb_agent_waypoints(lmb_ctx* ctx, uint32_t capacity, uint32_t* out_kind, float* out_points,
                            uint32_t* out_link, uint32_t* out_count);

/* `debug::draw_avoidance_data` (debug.rs:415-503, feature `debug-avoidance`): the velocity-space
 * constraint lines the last step's avoidance built for the agent at `agent_index` of that step's
 * arrays — what `agent.avoidance_data` holds after avoidance.rs:161-172.  Only agents uploaded with
 * LMB_AGENT_KEEP_AVOIDANCE_DATA have any.  out_lines = [capacity][4] {point.xy, direction.xy} in
 * dodgy_2d's Line convention (ConstraintLine.normal = (-direction.y, direction.x)):
 * *out_n_original lines of DebugData::{Satisfied.constraints, Fallback.original_constraints}
 * (obstacle lines first, then one per neighbour) followed by *out_n_fallback lines of
 * Fallback.fallback_constraints (the projected set of the last 3-D re-solve).  *out_ran = 1:
 * DebugData::Satisfied, 2: DebugData::Fallback, 0: the step skipped avoidance for this agent (not
 * sampled, paused ...: the reference then leaves agent.avoidance_data as it was).  LMB_ERR_CAPACITY with the counts set if capacity is too small. */
int32_t lmb_agent_avoidance_constraints(lmb_ctx* ctx, uint32_t agent_index, uint32_t capacity, float* out_lines,
                                        uint32_t* out_n_original, uint32_t* out_n_fallback, uint32_t* out_ran);

int32_t lmb_get_step_timings(lmb_ctx* ctx, lmb_step_timings* out);
/* The same figures for the last lmb_find_paths / lmb_find_straight_paths batch (astar_ms, n_repath = queries,
 * explored_nodes, path_nodes, kernel_launches); the batched queries leave the step's timings alone. */
int32_t lmb_get_query_timings(lmb_ctx* ctx, lmb_step_timings* out);

#ifdef __cplusplus
}
#endif
#endif /* LANDMASS_B200_H */
