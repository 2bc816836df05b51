// landmass_b200 — C ABI implementation (include/landmass_b200.h): context, navdata
// upload, and the orchestration of one `Archipelago::update` on the device.
//
// Step pipeline on one stream (reference lib.rs:270-534):
//   H2D inputs -> k_sample_points (agents, targets, characters)      phase A / A'
//              -> k_repath_decide (decision + repath queue)          phase B (decision)
//              -> [sync: n_repath] -> k_astar (loop on overflow)     phase B (A*)
//              -> k_follow                                           phase C
//              -> avoidance kernels                                  phase D
//              -> k_gather_outputs -> D2H outputs
// There is no CPU fallback anywhere in this file: every phase is a kernel.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "lmb_context.h"

static thread_local std::string g_create_error;

void lmb_comm_release(lmb_ctx* ctx);  // lmb_comm.cu
int32_t lmb_step_begin_impl(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first,
                            bool sync_at_end);

// NVTX range per phase: makes every nsys / ncu capture attributable to a phase of the step.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};


// ---------------------------------------------------------------------------
// host math mirroring the device (no FMA: compiled with -ffp-contract=off)
// ---------------------------------------------------------------------------
namespace {
struct H3 {
  float x, y, z;
};
inline H3 h_rotate_z(float s, float c, H3 v) {
  volatile float b2 = (0.0f * 0.0f + 0.0f * 0.0f) + s * s;
  volatile float k1 = c * c - b2;
  volatile float vb = (v.x * 0.0f + v.y * 0.0f) + v.z * s;
  volatile float k2 = vb * 2.0f;
  volatile float k3 = c * 2.0f;
  volatile float bx = 0.0f * v.z - v.y * s, by = s * v.x - v.z * 0.0f, bz = 0.0f * v.y - v.x * 0.0f;
  volatile float x = (v.x * k1 + 0.0f * k2) + bx * k3;
  volatile float y = (v.y * k1 + 0.0f * k2) + by * k3;
  volatile float z = (v.z * k1 + s * k2) + bz * k3;
  return {x, y, z};
}
inline H3 h_apply(const DevIsland& is, H3 p) {
  H3 r = h_rotate_z(is.fs, is.fc, p);
  volatile float x = r.x + is.tx, y = r.y + is.ty, z = r.z + is.tz;
  return {x, y, z};
}
}  // namespace

extern "C" {

uint32_t lmb_abi_version(void) { return LMB_ABI_VERSION; }

const char* lmb_last_error(const lmb_ctx* ctx) {
  if (!ctx) return g_create_error.c_str();
  return ctx->error.c_str();
}

int32_t lmb_create(const lmb_config* config, lmb_ctx** out_ctx) {
  if (!config || !out_ctx || config->struct_size != sizeof(lmb_config)) {
    g_create_error = "lmb_create: bad config (struct_size mismatch?)";
    return LMB_ERR_INVALID_ARGUMENT;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("lmb_create: no CUDA device available (") + cudaGetErrorString(e) +
                     "); landmass_b200 has no CPU fallback";
    return LMB_ERR_CUDA;
  }
  if (config->device < 0 || config->device >= count) {
    g_create_error = "lmb_create: device ordinal out of range";
    return LMB_ERR_INVALID_ARGUMENT;
  }
  lmb_ctx* ctx = new lmb_ctx();
  ctx->cfg = *config;
  ctx->device = config->device;
  if (ctx->cfg.max_agents == 0) ctx->cfg.max_agents = 1024;
  e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 11 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev_q[i]);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev_ag[i]);
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_counters, sizeof(AStarCounters));
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_small, 64 * sizeof(uint32_t));
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, ctx->device);
  if (e != cudaSuccess) {
    g_create_error = std::string("lmb_create: ") + cudaGetErrorString(e);
    delete ctx;
    return LMB_ERR_CUDA;
  }
  ctx->sm_count = (uint32_t)prop.multiProcessorCount;
  const uint32_t m = ctx->cfg.max_agents;
  cudaError_t es[] = {
      ctx->d_slot_path_off.reserve(m), ctx->d_slot_path_len.reserve(m), ctx->d_slot_cost_hint.reserve(m), ctx->d_scan_tmp.reserve(m + 1),
      ctx->d_desired_move.reserve(3 * (size_t)m), ctx->d_reached_start.reserve(3 * (size_t)m),
      ctx->d_reached_end.reserve(3 * (size_t)m), ctx->d_state.reserve(m),
      ctx->d_reached_link_id.reserve(m), ctx->d_path_endpoints.reserve(6 * (size_t)m), ctx->d_pool_cursor.reserve(1), ctx->d_counters.reserve(1),
      ctx->d_queue_count.reserve(1), ctx->d_avoid_overflow.reserve(4), ctx->d_bounds_tmp.reserve(8)};
  for (cudaError_t x : es)
    if (x != cudaSuccess) {
      g_create_error = std::string("lmb_create: ") + cudaGetErrorString(x);
      delete ctx;
      return LMB_ERR_CUDA;
    }
  cudaMemsetAsync(ctx->d_slot_path_len.p, 0, m * sizeof(uint32_t), ctx->stream);
  cudaMemsetAsync(ctx->d_slot_path_off.p, 0, m * sizeof(uint32_t), ctx->stream);
  cudaMemsetAsync(ctx->d_slot_cost_hint.p, 0, m * sizeof(uint32_t), ctx->stream);
  cudaMemsetAsync(ctx->d_desired_move.p, 0, 3 * (size_t)m * sizeof(float), ctx->stream);
  cudaMemsetAsync(ctx->d_state.p, 0, m * sizeof(uint32_t), ctx->stream);  // LMB_STATE_IDLE
  cudaMemsetAsync(ctx->d_reached_link_id.p, 0xff, m * sizeof(uint32_t), ctx->stream);
  cudaMemsetAsync(ctx->d_pool_cursor.p, 0, sizeof(unsigned long long), ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  *out_ctx = ctx;
  return LMB_OK;
}

void lmb_destroy(lmb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  for (int i = 0; i < 11; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; i++) {
    if (ctx->ev_q[i]) cudaEventDestroy(ctx->ev_q[i]);
    if (ctx->ev_ag[i]) cudaEventDestroy(ctx->ev_ag[i]);
  }
  lmb_comm_release(ctx);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  delete ctx;
}

// ---------------------------------------------------------------------------
// navdata upload
// ---------------------------------------------------------------------------
static PathPool make_pool(lmb_ctx* ctx) {
  PathPool p;
  p.entries = ctx->d_pool[ctx->pool_cur].p;
  p.cursor = ctx->d_pool_cursor.p;
  p.capacity = ctx->pool_capacity;
  p.slot_path_off = ctx->d_slot_path_off.p;
  p.slot_path_len = ctx->d_slot_path_len.p;
  p.slot_cost_hint = ctx->d_slot_cost_hint.p;
  return p;
}

static const uint32_t OVF_CAP = 1024;  // compact layout: overflow-table entries per slot (power of two)

// Byte layout of one arena slot: best_estimates region | node list | open-list spill |
// overflow table (compact layout only).
static void arena_layout(const lmb_ctx* ctx, uint32_t node_cap, uint32_t heap_cap, int layout, AStarArena& a) {
  a.n_ids = ctx->n_ids;
  a.kshift = ctx->kshift;
  a.n_states = ctx->n_ids + ctx->nav.n_links + 2;
  a.node_cap = node_cap;
  a.heap_cap = heap_cap;
  a.layout = (uint32_t)layout;
  a.ovf_cap = layout == 1 ? OVF_CAP : 0u;
  a.hash_cap = 0;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // the kernels address the spill as gm[i - HS + 1] with the HS of the launch shape, i < heap_cap: size it for
  // the smallest HS any shape uses (sizing it from AT_HEAP_SM = 48 let the <16,40,5> shape write 7
  // entries past the region)
  const size_t spill = (heap_cap > AT_HEAP_SM_MIN ? heap_cap - AT_HEAP_SM_MIN : 0) + 2;
  a.special_offset = 0;
  if (layout == 1) {
    const size_t n_polys = ctx->n_ids >> ctx->kshift;
    a.special_offset = up(n_polys * 8);
    a.best_bytes = a.special_offset + up((size_t)(ctx->nav.n_links + 2) * 4);
  } else if (layout == 2) {
    uint32_t cap = 1024;
    while (cap < 2u * node_cap + 16u) cap <<= 1;  // every node adds at most one key: never more than half full
    a.hash_cap = cap;
    a.best_bytes = (size_t)cap * 8;
  } else if (layout == 3) {
    a.best_bytes = 0;  // forest: no best_estimates at all
  } else {
    a.best_bytes = up((size_t)a.n_states * 4);
  }
  a.nodes_offset = a.best_bytes;
  a.heap_offset = a.nodes_offset + up((size_t)node_cap * 8);
  a.ovf_offset = a.heap_offset + up(spill * 16);
  a.slot_stride = a.ovf_offset + up((size_t)a.ovf_cap * 8);
}

// (Re)allocates an A* arena: `slots` query slots.
static int32_t setup_arena_into(lmb_ctx* ctx, AStarArena& a, DevBuf<uint8_t>& buf, uint32_t node_cap,
                                uint32_t heap_cap, uint32_t slots, int layout, const char* what) {
  arena_layout(ctx, node_cap, heap_cap, layout, a);
  a.n_slots = slots;
  a.base = nullptr;
  const size_t bytes = a.slot_stride * (size_t)slots;
  if (buf.cap < bytes) {
    buf.release();
    CUDA_TRY(ctx, buf.reserve_exact(bytes));
  }
  a.base = buf.p;
  if (ctx->cfg.flags & LMB_CONFIG_DEBUG_LOG)
    fprintf(stderr, "[lmb] A* %s: %u slots x %.2f MB (n_states %u, node_cap %u, heap_cap %u, kshift %u, %s best)\n",
            what, slots, a.slot_stride / 1048576.0, a.n_states, node_cap, heap_cap, a.kshift,
            a.layout == 1 ? "compact" : (a.layout == 2 ? "hashed" : (a.layout == 3 ? "no" : "dense")));
  launch_arena_init(a, ctx->stream);
  CUDA_TRY(ctx, cudaGetLastError());
  return LMB_OK;
}

static int32_t setup_arena(lmb_ctx* ctx, uint32_t node_cap, uint32_t heap_cap, uint32_t slots) {
  ctx->d_arena.release();
  return setup_arena_into(ctx, ctx->arena, ctx->d_arena, node_cap, heap_cap, slots, ctx->best_layout, "arena");
}

static size_t arena_slot_bytes(const lmb_ctx* ctx, uint32_t node_cap, uint32_t heap_cap) {
  AStarArena a{};
  arena_layout(ctx, node_cap, heap_cap, ctx->best_layout, a);
  return a.slot_stride;
}

static size_t arena_budget(const lmb_ctx* ctx) {
  if (ctx->cfg.scratch_bytes) return (size_t)ctx->cfg.scratch_bytes;
  // auto: what is free now plus what the current arena already holds, minus head-room for the
  // path pool (two copies during compaction) and the per-frame agent arrays
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)16 << 30;
  free_b += ctx->d_arena.cap;
  return std::min<size_t>((size_t)(free_b * 0.4), (size_t)64 << 30);
}

// Makes sure the arena can run `n_queries` concurrently (up to the slot limit / byte budget).
static int32_t ensure_arena(lmb_ctx* ctx, uint32_t n_queries) {
  const uint32_t max_slots = ctx->cfg.astar_slots ? ctx->cfg.astar_slots : ctx->sm_count * astar_resident_queries_per_sm(ctx->big_heap, ctx->best_layout == 3 ? ctx->cfg.astar_shape : 0u);
  uint32_t want = std::min<uint32_t>(max_slots, (n_queries + AT_THREADS - 1) / AT_THREADS * AT_THREADS);
  want = std::max<uint32_t>(want, AT_THREADS);
  if (ctx->arena.base && ctx->arena.n_slots >= want) return LMB_OK;
  const size_t per_slot = arena_slot_bytes(ctx, ctx->arena_node_cap, ctx->arena_heap_cap);
  const size_t budget = arena_budget(ctx);
  uint32_t slots = want;
  if (per_slot * slots > budget) slots = (uint32_t)std::max<size_t>(budget / per_slot, 32);
  if (ctx->arena.base && ctx->arena.n_slots >= slots) return LMB_OK;
  return setup_arena(ctx, ctx->arena_node_cap, ctx->arena_heap_cap, slots);
}

// Side arena (dense slots) of the small batches of a mesh whose main arena is hashed.  The hashed layout is for
// thousands of searches in flight on a mesh too large for a dense array each; the few hundred re-plans of a steady
// frame can afford 4 bytes per state and slot, and a dense probe is one plain load into an array whose touched
// part stays in L2 (1 M-polygon world, 100 000 agents: steady frame 291 -> 220 ms; forcing dense for every batch
// instead doubles frame 0, 7.6 -> 16.2 s).  Returns false when it does not fit the byte budget.
static const uint32_t SIDE_SLOTS_PER_SM = 4;
static bool ensure_side_arena(lmb_ctx* ctx, uint32_t n_queries, int32_t* status) {
  *status = LMB_OK;
  const uint32_t want = std::max<uint32_t>((n_queries + 31u) / 32u * 32u, 64u);
  if (!ctx->side_node_cap) {
    const uint32_t n_states = ctx->n_ids + ctx->nav.n_links + 2;
    ctx->side_node_cap = std::max<uint32_t>(ctx->arena_node_cap, std::min<uint32_t>(n_states / 8 + 1024, 1u << 20));
    ctx->side_heap_cap = ctx->arena_heap_cap;
  }
  AStarArena probe{};
  arena_layout(ctx, ctx->side_node_cap, ctx->side_heap_cap, 0, probe);
  if (ctx->side_arena.base && ctx->side_arena.n_slots >= want && ctx->side_arena.node_cap == ctx->side_node_cap &&
      ctx->side_arena.heap_cap == ctx->side_heap_cap)
    return true;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  free_b += ctx->d_side_arena.cap;
  const size_t budget = ctx->cfg.scratch_bytes ? (size_t)ctx->cfg.scratch_bytes / 4 : std::min<size_t>(free_b / 4, (size_t)16 << 30);
  const uint32_t slots = std::max(want, ctx->side_arena.base ? ctx->side_arena.n_slots : 0u);
  if (probe.slot_stride * (size_t)slots > budget) return false;
  ctx->d_side_arena.release();
  ctx->side_arena = AStarArena{};
  *status = setup_arena_into(ctx, ctx->side_arena, ctx->d_side_arena, ctx->side_node_cap, ctx->side_heap_cap, slots, 0,
                             "side arena");
  return *status == LMB_OK;
}

int32_t lmb_navdata_upload(lmb_ctx* ctx, const lmb_navdata_desc* d) { return lmb_navdata_update(ctx, d, nullptr); }

int32_t lmb_navdata_update(lmb_ctx* ctx, const lmb_navdata_desc* d, const lmb_nav_remap* remap) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!d || d->struct_size != sizeof(lmb_navdata_desc))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_navdata_upload: bad descriptor (struct_size)");
  if (remap && ctx->has_nav) {
    if (remap->struct_size != sizeof(lmb_nav_remap) || remap->n_old_islands + 1 != ctx->h_island_poly_begin.size() ||
        remap->n_old_links != ctx->nav.n_links || (remap->n_old_islands && !remap->island_map) ||
        (remap->n_old_links && !remap->link_map))
      return ctx->fail(LMB_ERR_INVALID_ARGUMENT,
                       "lmb_navdata_update: remap does not describe the previous upload (%u islands, %u links)",
                       (unsigned)(ctx->h_island_poly_begin.size() - 1), ctx->nav.n_links);
    for (uint32_t i = 0; i < remap->n_old_islands; i++) {
      const uint32_t j = remap->island_map[i];
      if (j == LMB_NONE) continue;
      if (j >= d->n_islands || d->island_poly_begin[j + 1] - d->island_poly_begin[j] !=
                                   ctx->h_island_poly_begin[i + 1] - ctx->h_island_poly_begin[i])
        return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_navdata_update: island_map[%u] = %u is not the same island", i, j);
    }
    for (uint32_t l = 0; l < remap->n_old_links; l++)
      if (remap->link_map[l] != LMB_NONE && remap->link_map[l] >= d->n_links)
        return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_navdata_update: link_map[%u] out of range", l);
  }
  const bool carry_paths = remap && ctx->has_nav;
  const std::vector<uint32_t> old_island_poly_begin = ctx->h_island_poly_begin;
  const uint32_t n_old_links = ctx->nav.n_links;
  ctx->has_nav = false;  // a failure below leaves the context without navigation data, not with half of it
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  cudaStreamSynchronize(s);
  DevNav& nav = ctx->nav;
  nav = DevNav{};
  nav.n_islands = d->n_islands;
  nav.n_polys = d->n_polys;
  nav.n_edges = d->n_edges;
  nav.n_vertices = d->n_vertices;
  nav.n_links = d->n_links;

  // ---- dense type ids ----
  std::map<uint32_t, uint32_t> type_dense;
  for (uint32_t p = 0; p < d->n_polys; p++) type_dense[d->poly_type[p]] = 0;
  for (uint32_t l = 0; l < d->n_links; l++) type_dense[d->link_dst_type[l]] = 0;
  for (uint32_t t = 0; t < d->n_type_costs; t++) type_dense[d->type_cost_index[t]] = 0;
  if (type_dense.empty()) type_dense[0] = 0;
  if (type_dense.size() > LMB_MAX_TYPES)
    return ctx->fail(LMB_ERR_UNSUPPORTED, "more than %u distinct node types", LMB_MAX_TYPES);
  std::vector<uint32_t> type_raw;
  for (auto& kv : type_dense) {
    kv.second = (uint32_t)type_raw.size();
    type_raw.push_back(kv.first);
  }
  nav.n_types = (uint32_t)type_raw.size();
  std::vector<float> type_cost(nav.n_types, 1.0f);
  std::vector<uint8_t> type_registered(nav.n_types, 0);
  for (uint32_t t = 0; t < d->n_type_costs; t++) {
    uint32_t k = type_dense[d->type_cost_index[t]];
    type_cost[k] = d->type_cost_value[t];
    type_registered[k] = 1;
  }

  // ---- islands, world vertices, sampling grids ----
  std::vector<DevIsland> islands(d->n_islands);
  std::vector<uint32_t> poly_island(d->n_polys), poly_type_dense(d->n_polys);
  std::vector<float> verts_world((size_t)d->n_vertices * 3);
  std::vector<uint32_t> cell_start, cell_polys;
  bool any = false;
  for (uint32_t i = 0; i < d->n_islands; i++) {
    DevIsland& is = islands[i];
    is.tx = d->island_translation[3 * i];
    is.ty = d->island_translation[3 * i + 1];
    is.tz = d->island_translation[3 * i + 2];
    float rot = d->island_rotation[i];
    volatile float hf = rot * 0.5f;
    volatile float hi = (-rot) * 0.5f;
    is.fs = sinf(hf);
    is.fc = cosf(hf);
    is.is = sinf(hi);
    is.ic = cosf(hi);
    for (int k = 0; k < 3; k++) {
      is.bmin[k] = d->island_bounds[6 * i + k];
      is.bmax[k] = d->island_bounds[6 * i + 3 + k];
    }
    is.bounds_empty = !(is.bmin[0] <= is.bmax[0]);
    is.poly_begin = d->island_poly_begin[i];
    is.poly_end = d->island_poly_begin[i + 1];
    is.vert_begin = d->island_vert_begin[i];
    is.vert_end = d->island_vert_begin[i + 1];
    is.has_height = d->island_has_height ? d->island_has_height[i] : 0;
    for (uint32_t p = is.poly_begin; p < is.poly_end; p++) poly_island[p] = i;
    for (uint32_t v = is.vert_begin; v < is.vert_end; v++) {
      H3 w = h_apply(is, {d->vertices[3 * (size_t)v], d->vertices[3 * (size_t)v + 1], d->vertices[3 * (size_t)v + 2]});
      verts_world[3 * (size_t)v] = w.x;
      verts_world[3 * (size_t)v + 1] = w.y;
      verts_world[3 * (size_t)v + 2] = w.z;
      if (!any) {
        ctx->world_min[0] = ctx->world_max[0] = w.x;
        ctx->world_min[1] = ctx->world_max[1] = w.y;
        any = true;
      } else {
        ctx->world_min[0] = std::min(ctx->world_min[0], w.x);
        ctx->world_max[0] = std::max(ctx->world_max[0], w.x);
        ctx->world_min[1] = std::min(ctx->world_min[1], w.y);
        ctx->world_max[1] = std::max(ctx->world_max[1], w.y);
      }
    }
    // grid over polygon AABBs: ~1 polygon per cell, capped
    uint32_t np = is.poly_end - is.poly_begin;
    float ex = is.bounds_empty ? 1.0f : std::max(is.bmax[0] - is.bmin[0], 1e-3f);
    float ey = is.bounds_empty ? 1.0f : std::max(is.bmax[1] - is.bmin[1], 1e-3f);
    double target_cells = std::min<double>(std::max<double>(np, 1.0), 4.0e6);
    float cell = (float)std::sqrt((double)ex * ey / target_cells);
    if (!(cell > 0.0f) || !std::isfinite(cell)) cell = 1.0f;
    is.gx0 = is.bounds_empty ? 0.0f : is.bmin[0];
    is.gy0 = is.bounds_empty ? 0.0f : is.bmin[1];
    is.ginv = 1.0f / cell;
    is.gnx = (uint32_t)std::min<double>(std::max<double>(std::ceil(ex / cell), 1.0), 4096.0);
    is.gny = (uint32_t)std::min<double>(std::max<double>(std::ceil(ey / cell), 1.0), 4096.0);
    is.cell_begin = (uint32_t)cell_start.size();
    size_t ncell = (size_t)is.gnx * is.gny;
    std::vector<uint32_t> counts(ncell + 1, 0);
    auto range = [&](uint32_t p, int32_t& x0, int32_t& x1, int32_t& y0, int32_t& y1) {
      const float* b = d->poly_bounds + 6 * (size_t)p;
      x0 = grid_cell(b[0], is.gx0, is.ginv, is.gnx);
      x1 = grid_cell(b[3], is.gx0, is.ginv, is.gnx);
      y0 = grid_cell(b[1], is.gy0, is.ginv, is.gny);
      y1 = grid_cell(b[4], is.gy0, is.ginv, is.gny);
    };
    for (uint32_t p = is.poly_begin; p < is.poly_end; p++) {
      const float* b = d->poly_bounds + 6 * (size_t)p;
      if (!(b[0] <= b[3])) continue;
      int32_t x0, x1, y0, y1;
      range(p, x0, x1, y0, y1);
      for (int32_t y = y0; y <= y1; y++)
        for (int32_t x = x0; x <= x1; x++) counts[(size_t)y * is.gnx + x + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) counts[c + 1] += counts[c];
    size_t base = cell_polys.size();
    cell_polys.resize(base + counts[ncell]);
    std::vector<uint32_t> cursor(counts.begin(), counts.end() - 1);
    for (uint32_t p = is.poly_begin; p < is.poly_end; p++) {  // ascending polygon order per cell
      const float* b = d->poly_bounds + 6 * (size_t)p;
      if (!(b[0] <= b[3])) continue;
      int32_t x0, x1, y0, y1;
      range(p, x0, x1, y0, y1);
      for (int32_t y = y0; y <= y1; y++)
        for (int32_t x = x0; x <= x1; x++) cell_polys[base + cursor[(size_t)y * is.gnx + x]++] = p;
    }
    for (size_t c = 0; c <= ncell; c++) cell_start.push_back((uint32_t)(base + counts[c]));
  }
  for (uint32_t p = 0; p < d->n_polys; p++) poly_type_dense[p] = type_dense[d->poly_type[p]];

  // ---- node -> links CSR (may be absent) ----
  std::vector<uint32_t> node_link_begin(d->n_polys + 1, 0), node_links;
  if (d->node_link_begin && d->n_links) {
    node_link_begin.assign(d->node_link_begin, d->node_link_begin + d->n_polys + 1);
    node_links.assign(d->node_links, d->node_links + d->n_links);
  }

  // ---- edge records (= A* states) ----
  // Record id space: padded (`poly << kshift | local edge`) when every polygon has <= 8 edges, so
  // the kernel can address a polygon's records from a state id alone; CSR edge ids otherwise.
  uint32_t max_ne = 0;
  for (uint32_t p = 0; p < d->n_polys; p++)
    max_ne = std::max(max_ne, d->poly_edge_begin[p + 1] - d->poly_edge_begin[p]);
  if (max_ne > 255) return ctx->fail(LMB_ERR_UNSUPPORTED, "a polygon has more than 255 edges");
  uint32_t kshift = max_ne <= 4 ? 2u : (max_ne <= 8 ? 3u : 0u);
  if (kshift && ((uint64_t)d->n_polys << kshift) >= 0xF0000000ull) kshift = 0;
  const uint32_t n_ids = kshift ? (d->n_polys << kshift) : d->n_edges;
  ctx->kshift = kshift;
  ctx->n_ids = n_ids;
  auto rec_id = [&](uint32_t p, uint32_t e) { return kshift ? ((p << kshift) | e) : d->poly_edge_begin[p] + e; };
  EdgeRec blank{};
  blank.twin = LMB_NONE;
  std::vector<EdgeRec> edges(std::max<uint32_t>(n_ids, 4u), blank);
  for (uint32_t p = 0; p < d->n_polys; p++) {
    uint32_t eb = d->poly_edge_begin[p], ne = d->poly_edge_begin[p + 1] - eb;
    const DevIsland& is = islands[poly_island[p]];
    bool has_links = node_link_begin[p + 1] > node_link_begin[p];
    const uint32_t n_rec = kshift ? (1u << kshift) : ne;
    for (uint32_t e = 0; e < n_rec; e++) {
      EdgeRec& r = edges[rec_id(p, e)];
      r.poly = p;
      r.first_edge = rec_id(p, 0);
      r.info = ne | (poly_type_dense[p] << 8) | (has_links ? (1u << 24) : 0u);
      if (e >= ne) continue;  // padding record: twin stays LMB_NONE
      uint32_t vi = d->edge_vertex[eb + (e == ne - 1 ? 0 : e + 1)];  // left  = vertex[e+1]
      uint32_t vj = d->edge_vertex[eb + e];                           // right = vertex[e]
      const float* a = d->vertices + 3 * (size_t)vi;
      const float* b = d->vertices + 3 * (size_t)vj;
      volatile float mx = (a[0] + b[0]) * 0.5f, my = (a[1] + b[1]) * 0.5f, mz = (a[2] + b[2]) * 0.5f;
      H3 w = h_apply(is, {mx, my, mz});
      r.mx = w.x;
      r.my = w.y;
      r.mz = w.z;
      uint32_t nb = d->edge_neighbor[eb + e];
      if (nb != LMB_NONE) {
        r.twin = rec_id(nb, d->edge_reverse[eb + e]);
        r.aux = d->edge_reverse[eb + e] & 0xffu;
        r.info |= poly_type_dense[nb] << 16;
      }
    }
  }

  // portal endpoints per edge record (world coordinates: the same verts_world values the funnel used to gather
  // through poly_edge_begin -> edge_vertex -> verts_world)
  std::vector<float4> portal_lr((size_t)std::max<uint32_t>(n_ids, 4u) * 2, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
  for (uint32_t p = 0; p < d->n_polys; p++) {
    const uint32_t eb = d->poly_edge_begin[p], ne = d->poly_edge_begin[p + 1] - eb;
    for (uint32_t e = 0; e < ne; e++) {
      const uint32_t li = d->edge_vertex[eb + (e == ne - 1 ? 0 : e + 1)], ri = d->edge_vertex[eb + e];
      const size_t k = (size_t)rec_id(p, e) * 2;
      portal_lr[k] = make_float4(verts_world[3 * (size_t)li], verts_world[3 * (size_t)li + 1], verts_world[3 * (size_t)li + 2], 0.0f);
      portal_lr[k + 1] = make_float4(verts_world[3 * (size_t)ri], verts_world[3 * (size_t)ri + 1], verts_world[3 * (size_t)ri + 2], 0.0f);
    }
  }
  std::vector<uint32_t> link_dst_type_dense(d->n_links);
  for (uint32_t l = 0; l < d->n_links; l++) link_dst_type_dense[l] = type_dense[d->link_dst_type[l]];
  std::vector<uint32_t> node_modified(d->n_polys, LMB_NONE);
  for (uint32_t m = 0; m < d->n_modified; m++) node_modified[d->modified_node[m]] = m;
  std::vector<uint32_t> node_region_number(d->n_polys, LMB_NONE);
  if (d->node_region_number) node_region_number.assign(d->node_region_number, d->node_region_number + d->n_polys);

#define UP(buf, ptr, n) CUDA_TRY(ctx, ctx->buf.upload(ptr, n, s))
#define UPV(buf, vec) CUDA_TRY(ctx, ctx->buf.upload(vec, s))
  UPV(d_islands, islands);
  UP(d_verts_local, d->vertices, (size_t)d->n_vertices * 3);
  UPV(d_verts_world, verts_world);
  UP(d_poly_edge_begin, d->poly_edge_begin, (size_t)d->n_polys + 1);
  UP(d_edge_vertex, d->edge_vertex, d->n_edges);
  UP(d_edge_neighbor, d->edge_neighbor, d->n_edges);
  UP(d_edge_reverse, d->edge_reverse, d->n_edges);
  UP(d_poly_bounds, d->poly_bounds, (size_t)d->n_polys * 6);
  UPV(d_poly_island, poly_island);
  UP(d_poly_region, d->poly_region, d->n_polys);
  UPV(d_poly_type_dense, poly_type_dense);
  if (d->hpoly) {
    UP(d_hpoly, d->hpoly, (size_t)d->n_polys * 4);
    UP(d_hverts, d->hverts, (size_t)d->n_hverts * 3);
    UP(d_htris, d->htris, (size_t)d->n_htris * 3);
  }
  UPV(d_cell_start, cell_start);
  UPV(d_cell_polys, cell_polys);
  const size_t edges_bytes = (edges.size() * sizeof(EdgeRec) + 255) & ~(size_t)255;
  const size_t dist_bytes = ctx->kshift ? ((size_t)ctx->n_ids << ctx->kshift) * sizeof(float) : 0;
  CUDA_TRY(ctx, ctx->d_astar_tables.reserve(edges_bytes + dist_bytes + 256));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_astar_tables.p, edges.data(), edges.size() * sizeof(EdgeRec), cudaMemcpyHostToDevice, s));
  UPV(d_portal_lr, portal_lr);
  UPV(d_type_cost, type_cost);
  UPV(d_type_registered, type_registered);
  UPV(d_type_raw, type_raw);
  UP(d_link_dst_node, d->link_dst_node, d->n_links);
  UPV(d_link_dst_type_dense, link_dst_type_dense);
  UP(d_link_portal, d->link_portal, (size_t)d->n_links * 6);
  UP(d_link_kind, d->link_kind, d->n_links);
  UP(d_link_reverse, d->link_reverse, d->n_links);
  UP(d_link_dst_portal, d->link_dst_portal ? d->link_dst_portal : d->link_portal, (size_t)d->n_links * 6);
  UP(d_link_cost, d->link_cost, d->n_links);
  UP(d_link_anim_kind, d->link_anim_kind, d->n_links);
  UP(d_link_anim_id, d->link_anim_id, d->n_links);
  UPV(d_node_link_begin, node_link_begin);
  UPV(d_node_links, node_links);
  UPV(d_node_modified, node_modified);
  UP(d_modified_edge_begin, d->modified_edge_begin, d->n_modified ? (size_t)d->n_modified + 1 : 0);
  UP(d_modified_edges, d->modified_edges, d->n_modified ? 2 * (size_t)d->modified_edge_begin[d->n_modified] : 0);
  UP(d_modified_vert_begin, d->modified_vert_begin, d->n_modified ? (size_t)d->n_modified + 1 : 0);
  UP(d_modified_verts, d->modified_verts, d->n_modified ? 2 * (size_t)d->modified_vert_begin[d->n_modified] : 0);
  UPV(d_node_region_number, node_region_number);
  UP(d_region_root, d->region_root, d->n_region_numbers);
  UP(d_possible_src, d->possible_link_src, d->n_possible_links);
  UP(d_possible_dst, d->possible_link_dst, d->n_possible_links);
  UP(d_possible_kind, d->possible_link_kind, d->n_possible_links);
#undef UP
#undef UPV
  nav.islands = ctx->d_islands.p;
  nav.verts_local = ctx->d_verts_local.p;
  nav.verts_world = ctx->d_verts_world.p;
  nav.poly_edge_begin = ctx->d_poly_edge_begin.p;
  nav.edge_vertex = ctx->d_edge_vertex.p;
  nav.edge_neighbor = ctx->d_edge_neighbor.p;
  nav.edge_reverse = ctx->d_edge_reverse.p;
  nav.poly_bounds = ctx->d_poly_bounds.p;
  nav.poly_island = ctx->d_poly_island.p;
  nav.poly_region = ctx->d_poly_region.p;
  nav.poly_type_dense = ctx->d_poly_type_dense.p;
  nav.hpoly = ctx->d_hpoly.p;
  nav.hverts = ctx->d_hverts.p;
  nav.htris = ctx->d_htris.p;
  nav.cell_start = ctx->d_cell_start.p;
  nav.cell_polys = ctx->d_cell_polys.p;
  nav.edges = reinterpret_cast<const EdgeRec*>(ctx->d_astar_tables.p);
  nav.portal_lr = ctx->d_portal_lr.p;
  nav.rec_kshift = ctx->kshift;
  nav.edge_dist = nullptr;
  if (ctx->kshift) {
    float* dist = reinterpret_cast<float*>(ctx->d_astar_tables.p + edges_bytes);
    launch_build_edge_dist(nav.edges, ctx->n_ids, ctx->kshift, dist, s);
    CUDA_TRY(ctx, cudaGetLastError());
    nav.edge_dist = dist;
  }
  {
    // L2 persistence for the tables (B200: 126 MB of L2; the tables are 12-200 MB).  Best effort: a device
    // without the feature, or a window larger than the limit, just keeps the default policy.
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    const size_t want = edges_bytes + dist_bytes;
    // Only when the tables fit the persisting carve-out whole (measured: config 2, 25 MB: +0.5 %; config 3's first
    // frame, 12 MB under a thrashing dense best_estimates: +4 %; config 4's 200 MB of tables with a partial
    // hitRatio: -20 %, the part of the window that does not persist is evicted first).
    if (max_persist > 0 && max_window > 0 && want <= (size_t)max_persist && want <= (size_t)max_window &&
        !(ctx->cfg.flags & LMB_CONFIG_NO_L2_WINDOW)) {
      const size_t carve = std::min<size_t>(want, (size_t)max_persist);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.base_ptr = ctx->d_astar_tables.p;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(want, (size_t)max_window);
      attr.accessPolicyWindow.hitRatio = want ? (float)std::min(1.0, (double)carve / (double)want) : 1.0f;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    } else {
      cudaStreamAttrValue attr{};  // num_bytes = 0: no window (drops the one a previous navdata may have set)
      cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  nav.type_cost = ctx->d_type_cost.p;
  nav.type_registered = ctx->d_type_registered.p;
  nav.link_dst_node = ctx->d_link_dst_node.p;
  nav.link_dst_type_dense = ctx->d_link_dst_type_dense.p;
  nav.link_portal = ctx->d_link_portal.p;
  nav.link_kind = ctx->d_link_kind.p;
  nav.link_reverse = ctx->d_link_reverse.p;
  nav.link_dst_portal = ctx->d_link_dst_portal.p;
  nav.link_cost = ctx->d_link_cost.p;
  nav.link_anim_kind = ctx->d_link_anim_kind.p;
  nav.link_anim_id = ctx->d_link_anim_id.p;
  nav.node_link_begin = ctx->d_node_link_begin.p;
  nav.node_links = ctx->d_node_links.p;
  nav.node_modified = ctx->d_node_modified.p;
  nav.modified_edge_begin = ctx->d_modified_edge_begin.p;
  nav.modified_edges = ctx->d_modified_edges.p;
  nav.modified_vert_begin = ctx->d_modified_vert_begin.p;
  nav.modified_verts = ctx->d_modified_verts.p;
  nav.node_region_number = ctx->d_node_region_number.p;
  nav.region_root = ctx->d_region_root.p;
  nav.n_possible_links = d->n_possible_links;
  nav.n_region_numbers = d->n_region_numbers;
  nav.possible_link_src = ctx->d_possible_src.p;
  nav.possible_link_dst = ctx->d_possible_dst.p;
  nav.possible_link_kind = ctx->d_possible_kind.p;

  // ---- A* arena: sized lazily by the first batch of queries (ensure_arena) ----
  {
    const uint32_t n_states = ctx->n_ids + nav.n_links + 2;
    ctx->arena_node_cap = std::max<uint32_t>(n_states / 4 + 1024, d->n_region_numbers + 64);
    ctx->arena_heap_cap = AT_HEAP_SM + std::min<uint32_t>(std::max<uint32_t>(1024, n_states / 64), 8192);
    // best_estimates layout: compact (one entry per polygon) when the polygon graph — edges and
    // off-mesh links — is a forest: a search then enters every polygon through one edge only.
    // With a single cycle a search comes round both ways and enters whole regions twice, so
    // anything else starts dense.  (A compact query that fills its overflow table is re-run
    // dense and switches the context for good, see run_astar; lmb_config.flags carries the test
    // hooks that force a starting layout.)
    {
      std::vector<uint32_t> parent(d->n_polys);
      for (uint32_t p = 0; p < d->n_polys; p++) parent[p] = p;
      auto find = [&](uint32_t x) {
        while (parent[x] != x) x = parent[x] = parent[parent[x]];
        return x;
      };
      uint64_t connections = 0, components = d->n_polys;
      auto join = [&](uint32_t a, uint32_t b) {
        connections++;
        a = find(a), b = find(b);
        if (a != b) parent[a] = b, components--;
      };
      for (uint32_t p = 0; p < d->n_polys; p++)
        for (uint32_t e = d->poly_edge_begin[p]; e < d->poly_edge_begin[p + 1]; e++)
          if (d->edge_neighbor[e] != LMB_NONE && d->edge_neighbor[e] > p) join(p, d->edge_neighbor[e]);
      std::vector<uint32_t> link_src(d->n_links, LMB_NONE);
      if (d->node_link_begin)
        for (uint32_t p = 0; p < d->n_polys; p++)
          for (uint32_t k = d->node_link_begin[p]; k < d->node_link_begin[p + 1]; k++) link_src[d->node_links[k]] = p;
      for (uint32_t l = 0; l < d->n_links; l++)
        if (link_src[l] != LMB_NONE) join(link_src[l], d->link_dst_node[l]);
      const bool forest = connections + components == d->n_polys;
      // polygons of the largest connected component: on a forest a search never creates more nodes than its
      // own component has polygons (+ Start / End), whatever else the world holds (e.g. one maze island per GPU)
      uint32_t max_component = 0;
      {
        std::vector<uint32_t> size(d->n_polys, 0);
        for (uint32_t p = 0; p < d->n_polys; p++) max_component = std::max(max_component, ++size[find(p)]);
      }
      const uint32_t n_states = ctx->n_ids + nav.n_links + 2;
      const uint32_t fl = ctx->cfg.flags;
      const bool f_compact = fl & LMB_CONFIG_ASTAR_COMPACT, f_dense = fl & LMB_CONFIG_ASTAR_DENSE;
      const bool f_hashed = fl & LMB_CONFIG_ASTAR_HASHED;
      ctx->nav_is_forest = forest;
      ctx->best_layout = 0;
      if (f_dense || f_hashed || f_compact) {
        // test hooks: force a layout (compact needs padded ids)
        ctx->best_layout = f_dense ? 0 : (f_compact && ctx->kshift != 0 ? 1 : (f_hashed ? 2 : 0));
      } else if (forest && !(fl & LMB_CONFIG_ASTAR_NO_FOREST)) {
        // Forest: the entered edge is excluded from a state's successors (pathfinding.rs:165-166), so a
        // search enters every polygon once and generates every state once — `best_estimates` can never
        // reject or make stale anything.  No table at all: layout 3.
        ctx->best_layout = 3;
      } else if (forest && ctx->kshift != 0) {
        ctx->best_layout = 1;
      } else if (n_states > (1u << 21)) {
        ctx->best_layout = 2;  // a search touches a small part of a mesh this large: keep only what it touches
      }
      if (ctx->best_layout == 2) ctx->arena_node_cap = std::min<uint32_t>(ctx->arena_node_cap, 131072u);
      if (f_hashed) ctx->arena_node_cap = std::min<uint32_t>(ctx->arena_node_cap, 2048u);
      // the are_nodes_connected BFS borrows the node list as seen[] / queue[] (2 x n_region_numbers words)
      ctx->arena_node_cap = std::max<uint32_t>(ctx->arena_node_cap, d->n_region_numbers + 64);
      if (ctx->best_layout == 3) {
        ctx->arena_node_cap = std::max<uint32_t>(max_component + 64, d->n_region_numbers + 64);
        ctx->arena_heap_cap = AT_HEAP_SM + 1024;  // a tree's open list is its unexpanded branches: small (doubles on overflow)
      }
      if (fl & LMB_CONFIG_ASTAR_SMALL_HEAP) ctx->arena_heap_cap = AT_HEAP_SM + 16;  // test hook: 64 entries
      ctx->big_heap = !forest;  // first guess: a wavefront; corrected from the first batch's open lists
    }
    ctx->arena = AStarArena{};
    ctx->d_arena.release();
    ctx->side_arena = AStarArena{};
    ctx->d_side_arena.release();
    ctx->side_node_cap = ctx->side_heap_cap = 0;
  }

  // ---- path pool: drop all paths ----
  if (ctx->pool_capacity == 0) {
    ctx->pool_capacity = std::max<unsigned long long>((unsigned long long)ctx->cfg.max_agents * 64ull, 1ull << 16);
    CUDA_TRY(ctx, ctx->d_pool[0].reserve_exact(ctx->pool_capacity));
    ctx->pool_cur = 0;
  }
  ctx->h_island_poly_begin.assign(d->island_poly_begin, d->island_poly_begin + d->n_islands + 1);
  DevBuf<uint32_t> d_old_begin, d_new_begin, d_link_map;
  if (carry_paths) {
    // Path::is_valid against (invalidated_off_mesh_links, invalidated_islands), then re-index
    std::vector<uint32_t> new_begin(remap->n_old_islands);
    for (uint32_t i = 0; i < remap->n_old_islands; i++)
      new_begin[i] = remap->island_map[i] == LMB_NONE ? LMB_NONE : d->island_poly_begin[remap->island_map[i]];
    CUDA_TRY(ctx, d_old_begin.upload(old_island_poly_begin.data(), old_island_poly_begin.size(), s));
    CUDA_TRY(ctx, d_new_begin.upload(new_begin.data(), new_begin.size(), s));
    CUDA_TRY(ctx, d_link_map.upload(remap->link_map, n_old_links, s));
    launch_remap_paths(make_pool(ctx), ctx->cfg.max_agents, d_old_begin.p, remap->n_old_islands, d_new_begin.p,
                       d_link_map.p, n_old_links, s);
  } else {
    launch_drop_all_paths(make_pool(ctx), ctx->cfg.max_agents, s);
  }
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_slot_cost_hint.p, 0, ctx->cfg.max_agents * sizeof(uint32_t), s));
  ctx->explored_estimate = 0.0;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  ctx->has_nav = true;
  return LMB_OK;
}

// ---------------------------------------------------------------------------
// A* driver shared by lmb_step and lmb_find_paths
// ---------------------------------------------------------------------------
// Path offsets are 32-bit: the pool never holds more than this many (node, portal) entries.
static const unsigned long long POOL_MAX_ENTRIES = 0xFFFF0000ull;

// Compacts the live corridors into a fresh buffer with room for `needed_free` more entries.
// Bump allocation leaves dead corridors behind (dropped / replaced paths); this is the only
// place they are reclaimed.  Exact-size allocation: the capacity does not creep.
static int32_t compact_pool(lmb_ctx* ctx, unsigned long long needed_free) {
  cudaStream_t s = ctx->stream;
  const uint32_t m = ctx->cfg.max_agents;
  int other = 1 - ctx->pool_cur;
  PathPool src = make_pool(ctx);
  // live size = sum of path lengths (the scan also yields every path's new offset)
  CUDA_TRY(ctx, ctx->d_scan_tiles.reserve(exclusive_scan_tmp_words(m)));
  launch_pool_scan(src, m, ctx->d_scan_tmp.p, ctx->d_scan_tiles.p, s);
  unsigned long long live = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&live, ctx->d_pool_cursor.p, sizeof(live), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  unsigned long long want = std::max<unsigned long long>(live + needed_free, 1ull << 16);
  if (want <= ctx->pool_capacity && live * 2 <= ctx->pool_capacity) want = ctx->pool_capacity;  // keep size
  if (want > POOL_MAX_ENTRIES)
    return ctx->fail(LMB_ERR_CAPACITY, "path pool would exceed %llu entries (live %llu + needed %llu)",
                     POOL_MAX_ENTRIES, live, needed_free);
  if (live == 0 && want <= ctx->pool_capacity) {
    // nothing to keep (every corridor was dropped or replaced): the scan already put the cursor at 0
    return LMB_OK;
  }
  // The two buffers ping-pong and stay allocated: freeing and re-allocating tens of GB every time
  // costs tens to hundreds of milliseconds of host time with the GPU idle.
  if (ctx->d_pool[other].cap < want) {
    ctx->d_pool[other].release();
    CUDA_TRY(ctx, ctx->d_pool[other].reserve_exact(want));
  }
  launch_pool_copy(src, ctx->d_pool[other].p, m, ctx->d_scan_tmp.p, s);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  ctx->pool_cur = other;
  ctx->pool_capacity = ctx->d_pool[other].cap;
  return LMB_OK;
}

// Enlarges the pool keeping every entry where it is (offsets stay valid).  Used by the batch
// query API, whose finished corridors are referenced by offset only — compaction, which keeps
// the corridors owned by agent slots, would drop them.
static int32_t grow_pool(lmb_ctx* ctx, unsigned long long needed_free) {
  cudaStream_t s = ctx->stream;
  int other = 1 - ctx->pool_cur;
  unsigned long long want = ctx->pool_capacity + needed_free;
  if (want > POOL_MAX_ENTRIES)
    return ctx->fail(LMB_ERR_CAPACITY, "path pool would exceed %llu entries", POOL_MAX_ENTRIES);
  ctx->d_pool[other].release();
  CUDA_TRY(ctx, ctx->d_pool[other].reserve_exact(want));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pool[other].p, ctx->d_pool[ctx->pool_cur].p,
                                ctx->pool_capacity * sizeof(uint2), cudaMemcpyDeviceToDevice, s));
  // allocations that failed still advanced the cursor: put it back at the old end
  const unsigned long long cursor = ctx->pool_capacity;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pool_cursor.p, &cursor, sizeof(cursor), cudaMemcpyHostToDevice, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  ctx->d_pool[ctx->pool_cur].release();
  ctx->pool_cur = other;
  ctx->pool_capacity = want;
  return LMB_OK;
}

// Makes room for the corridors n queries are expected to produce (compacting dead corridors away
// if needed).  A query that finishes its search but cannot allocate its corridor has to be
// searched again, so this runs before the searches; returns the entry count reserved.
static int32_t prepare_pool(lmb_ctx* ctx, uint32_t n, unsigned long long* needed_out) {
  cudaStream_t s = ctx->stream;
  double per_query = ctx->path_len_estimate > 0.0 ? ctx->path_len_estimate * 1.25 + 16.0
                                                  : std::min(std::max(2.0 * std::sqrt((double)ctx->nav.n_polys), 64.0), 65536.0);
  unsigned long long needed = (unsigned long long)(per_query * n) + 1024;
  needed = std::min(needed, POOL_MAX_ENTRIES / 2);  // first guess only; overflow doubles it
  unsigned long long cursor = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&cursor, ctx->d_pool_cursor.p, sizeof(cursor), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  if (cursor + needed > ctx->pool_capacity) {
    int32_t st = compact_pool(ctx, needed);
    if (st != LMB_OK) return st;
  }
  if (needed_out) *needed_out = needed;
  return LMB_OK;
}

static int32_t run_astar(lmb_ctx* ctx, AStarQueries q) {
  cudaStream_t s = ctx->stream;
  uint32_t n = q.n;
  CUDA_TRY(ctx, cudaMemsetAsync(q.out_status, 0, n * sizeof(uint32_t), s));
  unsigned long long needed = 0;
  {
    int32_t st = prepare_pool(ctx, n, &needed);
    if (st != LMB_OK) return st;
  }
  bool side = ctx->best_layout == 2 && n <= SIDE_SLOTS_PER_SM * ctx->sm_count &&
              !(ctx->cfg.flags & (LMB_CONFIG_ASTAR_HASHED | LMB_CONFIG_ASTAR_LARGE_SHAPES));
  if (side) {
    int32_t st = LMB_OK;
    side = ensure_side_arena(ctx, n, &st);
    if (st != LMB_OK) return st;
  }
  if (!side) {
    int32_t st = ensure_arena(ctx, n);
    if (st != LMB_OK) return st;
  }
  unsigned long long done_total = 0, nodes_total = 0, explored_total = 0, heap_max_total = 0;
  const uint32_t* list = nullptr;
  uint32_t n_run = n;
  if (!side && q.slot && n > ctx->arena.n_slots) {
    // more queries than can run at once: start the (probably) longest searches first
    CUDA_TRY(ctx, ctx->d_query_order.reserve(n));
    launch_order_queries(n, q.slot, ctx->d_slot_cost_hint.p, (uint32_t)std::max(ctx->explored_estimate, 1.0),
                         ctx->d_query_order.p, s);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->timings.kernel_launches += 1;
    list = ctx->d_query_order.p;
  }
  std::vector<uint32_t> h_status;
  const AStarArena* cur = side ? &ctx->side_arena : &ctx->arena;  // the arena of the next launch
  for (int iter = 0; iter < 24; iter++) {
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(AStarCounters), s));
    AStarQueries qq = q;
    qq.n = n_run;
    qq.list = list;
    ctx->timings.astar_launch_kind |= launch_astar(ctx->nav, *cur, qq, make_pool(ctx), ctx->d_counters.p,
                                                   ctx->d_type_raw.p, ctx->big_heap, ctx->cfg.flags,
                                                   cur->layout == 3 ? ctx->cfg.astar_shape : 0u, s);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, sizeof(AStarCounters),
                                  cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    ctx->timings.explored_nodes += ctx->h_counters->explored_total;
    ctx->timings.path_nodes += ctx->h_counters->path_nodes_total;
    ctx->timings.kernel_launches += 2;  // k_astar + k_path_fixup
    const uint32_t ra = ctx->h_counters->n_retry_arena, rh = ctx->h_counters->n_retry_heap;
    const uint32_t rp = ctx->h_counters->n_retry_pool, rd = ctx->h_counters->n_retry_dense;
    done_total += ctx->h_counters->n_done;
    nodes_total += ctx->h_counters->path_nodes_total;
    explored_total += ctx->h_counters->explored_total;
    if (done_total) ctx->path_len_estimate = (double)nodes_total / (double)done_total;
    if (done_total) ctx->explored_estimate = (double)explored_total / (double)done_total;
    heap_max_total += ctx->h_counters->heap_max_total;
    if (ctx->h_counters->n_failed)
      return ctx->fail(LMB_ERR_CAPACITY, "A*: a corridor did not fit the staging area (%u queries)", ctx->h_counters->n_failed);
    if (ra == 0 && rh == 0 && rp == 0 && rd == 0) {
      // launch shape of the next batch: open lists that outgrow shared memory want the big-heap shape
      if (done_total >= 16) ctx->big_heap = (double)heap_max_total / (double)done_total > 1.5 * AT_HEAP_SM;
      return LMB_OK;
    }
    // overflow: collect the retry list, grow what overflowed, run again
    h_status.resize(n);
    CUDA_TRY(ctx, cudaMemcpy(h_status.data(), q.out_status, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> retry;
    for (uint32_t i = 0; i < n; i++)
      if (h_status[i] >= 2) retry.push_back(i);
    if (rp) {
      needed *= 2;
      // agent paths survive compaction through their slots; batch-query corridors only by offset
      int32_t st = q.slot ? compact_pool(ctx, needed) : grow_pool(ctx, needed);
      if (st != LMB_OK) return st;
    }
    if (rd) ctx->best_layout = 0;  // compact best_estimates overflowed: dense from now on
    if (side && (ra || rh)) {
      // the side arena's node list / open-list spill: double what overflowed; back to the main arena if that
      // no longer fits
      if (ra) ctx->side_node_cap = cur->node_cap * 2;
      if (rh) ctx->side_heap_cap = AT_HEAP_SM + (cur->heap_cap - AT_HEAP_SM) * 2;
      int32_t st = LMB_OK;
      side = ensure_side_arena(ctx, n, &st);
      if (st != LMB_OK) return st;
      if (!side) {
        st = ensure_arena(ctx, n);
        if (st != LMB_OK) return st;
        cur = &ctx->arena;
      }
    } else if (ra || rh || rd) {
      // node list / hashed table (ra) or open-list spill (rh) too small: double what overflowed
      uint32_t node_cap = cur->node_cap, heap_cap = cur->heap_cap;
      if (ra) node_cap *= 2;
      if (rh) heap_cap = AT_HEAP_SM + (heap_cap - AT_HEAP_SM) * 2;
      if (ra && ctx->best_layout == 2 && !(ctx->cfg.flags & LMB_CONFIG_ASTAR_HASHED)) {
        // The hashed table holds 8 bytes for at least twice the node list; once that is no smaller than the dense
        // array (4 bytes per state) the searches are not local after all: dense from now on — smaller slots AND
        // one plain load per probe (1 M-polygon world, searches of 1.3-3.4 M nodes: 80-160 MB -> 25-49 MB per
        // slot, 1.74 -> 1.32 us per explored node alone, 2.74 -> 1.34 at one search per SM).
        AStarArena h{};
        arena_layout(ctx, std::max(ctx->arena_node_cap, node_cap), std::max(ctx->arena_heap_cap, heap_cap), 2, h);
        if (h.best_bytes >= (size_t)h.n_states * 4) ctx->best_layout = 0;
      }
      // Resize the arena for good (the slot count follows the byte budget).  Measured alternative,
      // rejected: a side arena of fat slots for the few searches that overflow, keeping the lean
      // main arena — the same searches overflow again next frame and are then searched twice.
      ctx->arena_node_cap = std::max(ctx->arena_node_cap, node_cap);
      ctx->arena_heap_cap = std::max(ctx->arena_heap_cap, heap_cap);
      const size_t per_slot = arena_slot_bytes(ctx, ctx->arena_node_cap, ctx->arena_heap_cap);
      const size_t budget = arena_budget(ctx);
      uint32_t slots = ctx->arena.n_slots;
      if (per_slot * slots > budget) slots = (uint32_t)std::max<size_t>(budget / per_slot, 8);
      int32_t st = setup_arena(ctx, ctx->arena_node_cap, ctx->arena_heap_cap, slots);
      if (st != LMB_OK) return st;
      cur = &ctx->arena;
    }
    CUDA_TRY(ctx, ctx->d_retry_list.upload(retry, s));
    list = ctx->d_retry_list.p;
    n_run = (uint32_t)retry.size();
  }
  return ctx->fail(LMB_ERR_CAPACITY, "A* scratch kept overflowing");
}

// A* of the batched query API: the last step's timings / PathingResult bookkeeping (ctx->timings) must not be
// disturbed by a query made between two steps; the batch's own figures go to ctx->query_timings.
static int32_t run_astar_query(lmb_ctx* ctx, const AStarQueries& q) {
  const lmb_step_timings saved = ctx->timings;
  ctx->timings = lmb_step_timings{};
  cudaEventRecord(ctx->ev_q[0], ctx->stream);
  const int32_t st = run_astar(ctx, q);
  cudaEventRecord(ctx->ev_q[1], ctx->stream);
  if (st == LMB_OK && cudaEventSynchronize(ctx->ev_q[1]) == cudaSuccess)
    cudaEventElapsedTime(&ctx->timings.astar_ms, ctx->ev_q[0], ctx->ev_q[1]);
  ctx->timings.n_repath = q.n;
  ctx->query_timings = ctx->timings;
  ctx->timings = saved;
  return st;
}

// ---------------------------------------------------------------------------
// batched queries
// ---------------------------------------------------------------------------
int32_t lmb_sample_points(lmb_ctx* ctx, uint32_t n, const float* points, const lmb_options* options,
                          float* out_points, uint32_t* out_nodes) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->has_nav) return ctx->fail(LMB_ERR_NO_NAVDATA, "no navigation data uploaded");
  if (!options || (n && (!points || !out_points || !out_nodes)))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_sample_points: null argument");
  if (n == 0) return LMB_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  // scratch of its own: the resident agent inputs (d_position, d_agent_point ...) must survive a query made
  // between two lmb_step_resident calls
  CUDA_TRY(ctx, ctx->d_query_point.upload(points, 3 * (size_t)n, s));
  CUDA_TRY(ctx, ctx->d_query_out_point.reserve(3 * (size_t)n));
  CUDA_TRY(ctx, ctx->d_query_out_node.reserve(n));
  launch_sample_points(ctx->nav, *options, n, ctx->d_query_point.p, nullptr, 0, 0, ctx->d_query_out_point.p,
                       ctx->d_query_out_node.p, s);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(out_points, ctx->d_query_out_point.p, 3 * (size_t)n * sizeof(float),
                                cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_nodes, ctx->d_query_out_node.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return LMB_OK;
}

static int32_t reserve_queries(lmb_ctx* ctx, uint32_t n) {
  size_t m = n ? n : 1;
  CUDA_TRY(ctx, ctx->d_q_agent.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_slot.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_start_node.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_end_node.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_start_point.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_q_end_point.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_q_status.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_success.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_explored.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_path_off.reserve(m));
  CUDA_TRY(ctx, ctx->d_q_path_len.reserve(m));
  return LMB_OK;
}

static AStarQueries make_queries(lmb_ctx* ctx, uint32_t n) {
  AStarQueries q{};
  q.n = n;
  q.start_node = ctx->d_q_start_node.p;
  q.end_node = ctx->d_q_end_node.p;
  q.start_point = ctx->d_q_start_point.p;
  q.end_point = ctx->d_q_end_point.p;
  q.out_status = ctx->d_q_status.p;
  q.out_success = ctx->d_q_success.p;
  q.out_explored = ctx->d_q_explored.p;
  q.out_path_off = ctx->d_q_path_off.p;
  q.out_path_len = ctx->d_q_path_len.p;
  return q;
}

// override_type_index_costs / permitted_animation_links of the batched queries (query.rs:185-191).
// They go to buffers of their own: the agents' copies (lmb_agents_upload) must survive a query
// made between two resident steps.  permit_all[i] plays the part of the agent flag
// LMB_AGENT_PERMIT_ALL_LINKS (AStarQueries::owner_flags).
static int32_t upload_query_filters(lmb_ctx* ctx, AStarQueries& q, uint32_t n, const uint32_t* override_begin,
                                    const uint32_t* override_type, const float* override_cost,
                                    const uint32_t* permit_all, const uint32_t* permitted_begin,
                                    const uint32_t* permitted_kind) {
  cudaStream_t s = ctx->stream;
  if (override_begin) {
    if (override_begin[n] && (!override_type || !override_cost))
      return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "override arrays are null but the CSR is not empty");
    CUDA_TRY(ctx, ctx->d_qf_override_begin.upload(override_begin, (size_t)n + 1, s));
    CUDA_TRY(ctx, ctx->d_qf_override_type.upload(override_type, override_begin[n], s));
    CUDA_TRY(ctx, ctx->d_qf_override_cost.upload(override_cost, override_begin[n], s));
    q.override_begin = ctx->d_qf_override_begin.p;
    q.override_type = ctx->d_qf_override_type.p;
    q.override_cost = ctx->d_qf_override_cost.p;
  }
  if (!permit_all) return LMB_OK;  // PermittedAnimationLinks::All for every query
  std::vector<uint32_t> flags(n), begin(n + 1, 0);
  for (uint32_t i = 0; i < n; i++) flags[i] = permit_all[i] ? LMB_AGENT_PERMIT_ALL_LINKS : 0u;
  if (permitted_begin) begin.assign(permitted_begin, permitted_begin + n + 1);
  if (begin[n] && !permitted_kind)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "permitted_kind is null but the CSR is not empty");
  CUDA_TRY(ctx, ctx->d_qf_flags.upload(flags.data(), n, s));
  CUDA_TRY(ctx, ctx->d_qf_permitted_begin.upload(begin.data(), (size_t)n + 1, s));
  CUDA_TRY(ctx, ctx->d_qf_permitted_kind.upload(permitted_kind, begin[n], s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));  // flags / begin are locals
  q.owner_flags = ctx->d_qf_flags.p;
  q.permitted_begin = ctx->d_qf_permitted_begin.p;
  q.permitted_kind = ctx->d_qf_permitted_kind.p;
  return LMB_OK;
}

int32_t lmb_find_paths(lmb_ctx* ctx, uint32_t n, const uint32_t* start_node, const float* start_point,
                       const uint32_t* end_node, const float* end_point,
                       const uint32_t* override_begin, const uint32_t* override_type,
                       const float* override_cost, uint32_t* out_success, uint32_t* out_explored,
                       uint32_t* out_path_begin, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t path_capacity, const uint32_t* permit_all, const uint32_t* permitted_begin,
                       const uint32_t* permitted_kind) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->has_nav) return ctx->fail(LMB_ERR_NO_NAVDATA, "no navigation data uploaded");
  if (!out_path_begin) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_paths: null argument");
  out_path_begin[0] = 0;
  if (n == 0) return LMB_OK;
  if (!start_node || !start_point || !end_node || !end_point || !out_success || !out_explored)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_paths: null argument");
  for (uint32_t i = 0; i < n; i++)
    if (start_node[i] >= ctx->nav.n_polys || end_node[i] >= ctx->nav.n_polys)
      return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_paths: node id out of range (query %u)", i);
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  int32_t st = reserve_queries(ctx, n);
  if (st != LMB_OK) return st;
  CUDA_TRY(ctx, ctx->d_q_start_node.upload(start_node, n, s));
  CUDA_TRY(ctx, ctx->d_q_end_node.upload(end_node, n, s));
  CUDA_TRY(ctx, ctx->d_q_start_point.upload(start_point, 3 * (size_t)n, s));
  CUDA_TRY(ctx, ctx->d_q_end_point.upload(end_point, 3 * (size_t)n, s));
  AStarQueries q = make_queries(ctx, n);
  st = upload_query_filters(ctx, q, n, override_begin, override_type, override_cost, permit_all, permitted_begin,
                            permitted_kind);
  if (st != LMB_OK) return st;
  // Paths of batch queries live in the pool only until they are copied out: remember the
  // cursor and rewind afterwards (agent paths are never touched).
  st = prepare_pool(ctx, n, nullptr);  // may compact: do it before the cursor is remembered
  if (st != LMB_OK) return st;
  unsigned long long cursor0 = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&cursor0, ctx->d_pool_cursor.p, sizeof(cursor0), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  st = run_astar_query(ctx, q);
  if (st != LMB_OK) return st;
  std::vector<uint32_t> off(n), len(n);
  CUDA_TRY(ctx, cudaMemcpyAsync(out_success, q.out_success, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(out_explored, q.out_explored, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(off.data(), q.out_path_off, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(len.data(), q.out_path_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) {
    out_path_begin[i] = (uint32_t)total;
    total += out_success[i] ? len[i] : 0;
  }
  out_path_begin[n] = (uint32_t)total;
  int32_t result = LMB_OK;
  if (total > path_capacity) {
    result = ctx->fail(LMB_ERR_CAPACITY, "lmb_find_paths: %llu path entries do not fit capacity %u",
                       (unsigned long long)total, path_capacity);
  } else if (total) {
    if (!out_nodes || !out_portals) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_paths: null outputs");
    std::vector<uint2> tmp;
    const uint2* pool = ctx->d_pool[ctx->pool_cur].p;
    for (uint32_t i = 0; i < n; i++) {
      if (!out_success[i] || !len[i]) continue;
      tmp.resize(len[i]);
      CUDA_TRY(ctx, cudaMemcpy(tmp.data(), pool + off[i], len[i] * sizeof(uint2), cudaMemcpyDeviceToHost));
      for (uint32_t k = 0; k < len[i]; k++) {
        out_nodes[out_path_begin[i] + k] = tmp[k].x;
        out_portals[out_path_begin[i] + k] = tmp[k].y;
      }
    }
  }
  // rewind the bump pointer only if no agent path was appended meanwhile (single-threaded use)
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pool_cursor.p, &cursor0, sizeof(cursor0), cudaMemcpyHostToDevice, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return result;
}

int32_t lmb_find_straight_paths(lmb_ctx* ctx, uint32_t n, const uint32_t* start_node,
                                const float* start_point, const uint32_t* end_node,
                                const float* end_point, const uint32_t* override_begin,
                                const uint32_t* override_type, const float* override_cost,
                                uint32_t* out_success, uint32_t* out_step_begin, uint32_t* out_kind,
                                float* out_points, uint32_t* out_link, uint32_t step_capacity,
                                const uint32_t* permit_all, const uint32_t* permitted_begin,
                                const uint32_t* permitted_kind) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->has_nav) return ctx->fail(LMB_ERR_NO_NAVDATA, "no navigation data uploaded");
  if (!out_step_begin) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_straight_paths: null argument");
  out_step_begin[0] = 0;
  if (n == 0) return LMB_OK;
  if (!start_node || !start_point || !end_node || !end_point || !out_success)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_straight_paths: null argument");
  for (uint32_t i = 0; i < n; i++)
    if (start_node[i] >= ctx->nav.n_polys || end_node[i] >= ctx->nav.n_polys)
      return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_straight_paths: node id out of range (query %u)", i);
  if (override_begin)
    for (uint32_t k = 0; k < override_begin[n]; k++)
      if (!(override_cost[k] > 0.0f))  // FindPathError::NonPositiveTypeIndexCost (query.rs:204-208)
        return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "non-positive type index cost override %u -> %f",
                         override_type[k], override_cost[k]);
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  int32_t st = reserve_queries(ctx, n);
  if (st != LMB_OK) return st;
  CUDA_TRY(ctx, ctx->d_q_start_node.upload(start_node, n, s));
  CUDA_TRY(ctx, ctx->d_q_end_node.upload(end_node, n, s));
  CUDA_TRY(ctx, ctx->d_q_start_point.upload(start_point, 3 * (size_t)n, s));
  CUDA_TRY(ctx, ctx->d_q_end_point.upload(end_point, 3 * (size_t)n, s));
  AStarQueries q = make_queries(ctx, n);
  st = upload_query_filters(ctx, q, n, override_begin, override_type, override_cost, permit_all, permitted_begin,
                            permitted_kind);
  if (st != LMB_OK) return st;
  st = prepare_pool(ctx, n, nullptr);  // may compact: do it before the cursor is remembered
  if (st != LMB_OK) return st;
  unsigned long long cursor0 = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&cursor0, ctx->d_pool_cursor.p, sizeof(cursor0), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  st = run_astar_query(ctx, q);
  if (st != LMB_OK) return st;
  std::vector<uint32_t> len(n), begin(n + 1), count(n);
  CUDA_TRY(ctx, cudaMemcpyAsync(out_success, q.out_success, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(len.data(), q.out_path_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  uint64_t total_cap = 0;
  for (uint32_t i = 0; i < n; i++) {
    begin[i] = (uint32_t)total_cap;
    total_cap += out_success[i] ? 2ull * len[i] + 2 : 0;  // generous bound on the number of steps
  }
  begin[n] = (uint32_t)total_cap;
  DevBuf<uint32_t> d_begin, d_count, d_kind, d_link;
  DevBuf<float> d_points;
  CUDA_TRY(ctx, d_begin.upload(begin, s));
  CUDA_TRY(ctx, d_count.reserve(n));
  CUDA_TRY(ctx, d_kind.reserve(total_cap + 1));
  CUDA_TRY(ctx, d_link.reserve(total_cap + 1));
  CUDA_TRY(ctx, d_points.reserve(6 * total_cap + 6));
  launch_straight_paths(ctx->nav, make_pool(ctx), q, d_begin.p, d_count.p, d_kind.p, d_points.p, d_link.p, s);
  CUDA_TRY(ctx, cudaGetLastError());
  std::vector<uint32_t> kind(total_cap + 1), link(total_cap + 1);
  std::vector<float> points(6 * total_cap + 6);
  CUDA_TRY(ctx, cudaMemcpyAsync(count.data(), d_count.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(kind.data(), d_kind.p, total_cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(link.data(), d_link.p, total_cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(points.data(), d_points.p, 6 * total_cap * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_pool_cursor.p, &cursor0, sizeof(cursor0), cudaMemcpyHostToDevice, s));
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) {
    out_step_begin[i] = (uint32_t)total;
    if (count[i] > begin[i + 1] - begin[i]) return ctx->fail(LMB_ERR_INTERNAL, "straight path of query %u truncated", i);
    total += count[i];
  }
  out_step_begin[n] = (uint32_t)total;
  if (total > step_capacity)
    return ctx->fail(LMB_ERR_CAPACITY, "lmb_find_straight_paths: %llu steps do not fit capacity %u",
                     (unsigned long long)total, step_capacity);
  if (total && (!out_kind || !out_points || !out_link))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_find_straight_paths: null outputs");
  for (uint32_t i = 0; i < n; i++)
    for (uint32_t k = 0; k < count[i]; k++) {
      uint32_t src = begin[i] + k, dst = out_step_begin[i] + k;
      out_kind[dst] = kind[src];
      out_link[dst] = link[src];
      for (int c = 0; c < 6; c++) out_points[6 * (size_t)dst + c] = points[6 * (size_t)src + c];
    }
  return LMB_OK;
}

// ---------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------
int32_t lmb_agents_upload(lmb_ctx* ctx, const lmb_agents_in* a, const lmb_characters_in* c) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!a) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_upload: null agents");
  const uint32_t n = a->n;
  if (n > ctx->cfg.max_agents)
    return ctx->fail(LMB_ERR_CAPACITY, "%u agents exceed max_agents %u", n, ctx->cfg.max_agents);
  if (n && (!a->slot || !a->flags || !a->position || !a->velocity || !a->radius || !a->desired_speed ||
            !a->max_speed))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_upload: null required array");
  for (uint32_t i = 0; i < n; i++)
    if (a->slot[i] >= ctx->cfg.max_agents)
      return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "agent %u: slot %u >= max_agents %u", i, a->slot[i],
                       ctx->cfg.max_agents);
  if (n && a->override_begin && a->override_begin[n] && (!a->override_type || !a->override_cost))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_upload: override arrays are null but the CSR is not empty");
  if (n && a->permitted_begin && a->permitted_begin[n] && !a->permitted_kind)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_upload: permitted_kind is null but the CSR is not empty");
  if (c && c->n && (!c->position || !c->velocity || !c->radius))
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_upload: null character array");
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  cudaEventRecord(ctx->ev[0], s);
  ctx->n_agents = n;
  CUDA_TRY(ctx, ctx->d_slot.upload(a->slot, n, s));
  CUDA_TRY(ctx, ctx->d_flags.upload(a->flags, n, s));
  // rows of the debug-avoidance export, in agent order
  ctx->dbg_rows = 0;
  ctx->dbg_lines_per_row = 0;  // no avoidance has run for these agents yet: nothing to report
  ctx->h_dbg_row_of.clear();
  for (uint32_t i = 0; i < n; i++) {
    if (!(a->flags[i] & LMB_AGENT_KEEP_AVOIDANCE_DATA)) continue;
    if (ctx->h_dbg_row_of.empty()) ctx->h_dbg_row_of.assign(n, LMB_NONE);
    ctx->h_dbg_row_of[i] = ctx->dbg_rows++;
  }
  if (ctx->dbg_rows) CUDA_TRY(ctx, ctx->d_dbg_row_of.upload(ctx->h_dbg_row_of, s));
  CUDA_TRY(ctx, ctx->d_position.upload(a->position, 3 * (size_t)n, s));
  CUDA_TRY(ctx, ctx->d_velocity.upload(a->velocity, 3 * (size_t)n, s));
  if (a->target) {
    CUDA_TRY(ctx, ctx->d_target.upload(a->target, 3 * (size_t)n, s));
  } else {
    CUDA_TRY(ctx, ctx->d_target.reserve(3 * (size_t)n + 1));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_target.p, 0, 3 * (size_t)n * sizeof(float), s));
  }
  CUDA_TRY(ctx, ctx->d_radius.upload(a->radius, n, s));
  CUDA_TRY(ctx, ctx->d_desired_speed.upload(a->desired_speed, n, s));
  CUDA_TRY(ctx, ctx->d_max_speed.upload(a->max_speed, n, s));
  ctx->has_reach_condition = a->reach_condition != nullptr;
  if (a->reach_condition) CUDA_TRY(ctx, ctx->d_reach_condition.upload(a->reach_condition, n, s));
  ctx->has_reach_distance = a->reach_distance != nullptr;
  if (a->reach_distance) CUDA_TRY(ctx, ctx->d_reach_distance.upload(a->reach_distance, n, s));
  ctx->has_anim_reach = a->anim_link_reach_distance != nullptr;
  if (a->anim_link_reach_distance) CUDA_TRY(ctx, ctx->d_anim_reach.upload(a->anim_link_reach_distance, n, s));
  ctx->has_overrides = a->override_begin != nullptr && n > 0 && a->override_begin[n] > 0;
  if (ctx->has_overrides) {
    CUDA_TRY(ctx, ctx->d_override_begin.upload(a->override_begin, (size_t)n + 1, s));
    CUDA_TRY(ctx, ctx->d_override_type.upload(a->override_type, a->override_begin[n], s));
    CUDA_TRY(ctx, ctx->d_override_cost.upload(a->override_cost, a->override_begin[n], s));
  }
  ctx->has_permitted = a->permitted_begin != nullptr && n > 0;
  if (ctx->has_permitted) {
    CUDA_TRY(ctx, ctx->d_permitted_begin.upload(a->permitted_begin, (size_t)n + 1, s));
    CUDA_TRY(ctx, ctx->d_permitted_kind.upload(a->permitted_kind, a->permitted_begin[n], s));
  }
  ctx->n_chars = c ? c->n : 0;
  if (ctx->n_chars) {
    CUDA_TRY(ctx, ctx->d_char_position.upload(c->position, 3 * (size_t)c->n, s));
    CUDA_TRY(ctx, ctx->d_char_velocity.upload(c->velocity, 3 * (size_t)c->n, s));
    CUDA_TRY(ctx, ctx->d_char_radius.upload(c->radius, c->n, s));
  }
  cudaEventRecord(ctx->ev[1], s);
  return LMB_OK;
}

} // extern C
AgentArrays lmb_make_agent_arrays(lmb_ctx* ctx) {
  AgentArrays ag{};
  ag.n = ctx->n_agents;
  ag.slot = ctx->d_slot.p;
  ag.flags = ctx->d_flags.p;
  ag.position = ctx->d_position.p;
  ag.velocity = ctx->d_velocity.p;
  ag.target = ctx->d_target.p;
  ag.radius = ctx->d_radius.p;
  ag.desired_speed = ctx->d_desired_speed.p;
  ag.max_speed = ctx->d_max_speed.p;
  ag.reach_condition = ctx->has_reach_condition ? ctx->d_reach_condition.p : nullptr;
  ag.reach_distance = ctx->has_reach_distance ? ctx->d_reach_distance.p : nullptr;
  ag.anim_link_reach_distance = ctx->has_anim_reach ? ctx->d_anim_reach.p : nullptr;
  ag.agent_point = ctx->d_agent_point.p;
  ag.agent_node = ctx->d_agent_node.p;
  ag.target_point = ctx->d_target_point.p;
  ag.target_node = ctx->d_target_node.p;
  return ag;
}

AgentPersist lmb_make_persist(lmb_ctx* ctx) {
  AgentPersist p{};
  p.desired_move = ctx->d_desired_move.p;
  p.state = ctx->d_state.p;
  p.reached_link_id = ctx->d_reached_link_id.p;
  p.reached_link_start = ctx->d_reached_start.p;
  p.reached_link_end = ctx->d_reached_end.p;
  p.path_endpoints = ctx->d_path_endpoints.p;
  return p;
}

// lmb_avoid_host.cu
int32_t lmb_avoidance_build_records(lmb_ctx* ctx, const lmb_options* o, uint32_t n_total, uint32_t first);
int32_t lmb_avoidance_solve(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first);

extern "C" {

int32_t lmb_step_begin(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first) {
  return lmb_step_begin_impl(ctx, o, dt, n_total, first, /*sync_at_end=*/true);
}

}  // extern "C"

// sync_at_end = false (lmb_step_sharded): the exchange and phase D are enqueued behind this on the same
// stream, so nothing has to wait for the host here.
int32_t lmb_step_begin_impl(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first,
                            bool sync_at_end) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->has_nav) return ctx->fail(LMB_ERR_NO_NAVDATA, "no navigation data uploaded");
  if (!o) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step: null options");
  if (first > n_total || ctx->n_agents > n_total - first)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step_begin: agent range [%u,+%u) outside %u", first,
                     ctx->n_agents, n_total);
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  const uint32_t n = ctx->n_agents;
  lmb_step_timings& T = ctx->timings;
  float h2d = 0.0f;
  if (cudaEventQuery(ctx->ev[1]) != cudaErrorInvalidResourceHandle) {
    cudaEventSynchronize(ctx->ev[1]);
    if (cudaEventElapsedTime(&h2d, ctx->ev[0], ctx->ev[1]) != cudaSuccess) h2d = 0.0f;
    cudaGetLastError();
  }
  T = lmb_step_timings{};
  T.h2d_ms = h2d;
  size_t m = n ? n : 1;
  CUDA_TRY(ctx, ctx->d_agent_point.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_agent_node.reserve(m));
  CUDA_TRY(ctx, ctx->d_target_point.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_target_node.reserve(m));
  CUDA_TRY(ctx, ctx->d_follow_start.reserve(m));
  CUDA_TRY(ctx, ctx->d_follow_end.reserve(m));
  CUDA_TRY(ctx, ctx->d_presult.reserve(m));
  CUDA_TRY(ctx, ctx->d_query_of_agent.reserve(m));
  CUDA_TRY(ctx, ctx->d_waypoint_kind.reserve(m));
  CUDA_TRY(ctx, ctx->d_waypoint_link.reserve(m));
  CUDA_TRY(ctx, ctx->d_waypoint_points.reserve(6 * m));
  CUDA_TRY(ctx, ctx->d_out_velocity.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_out_state.reserve(m));
  CUDA_TRY(ctx, ctx->d_out_link_id.reserve(m));
  CUDA_TRY(ctx, ctx->d_out_link_start.reserve(3 * m));
  CUDA_TRY(ctx, ctx->d_out_link_end.reserve(3 * m));
  int32_t st = reserve_queries(ctx, n);
  if (st != LMB_OK) return st;
  if (ctx->n_chars) {
    CUDA_TRY(ctx, ctx->d_char_point.reserve(3 * (size_t)ctx->n_chars));
    CUDA_TRY(ctx, ctx->d_char_node.reserve(ctx->n_chars));
  }
  AgentArrays ag = lmb_make_agent_arrays(ctx);
  AgentPersist persist = lmb_make_persist(ctx);

  NvtxRange nvtx_step("lmb:step_begin");
  cudaEventRecord(ctx->ev[2], s);
  nvtxRangePushA("lmb:phaseA sample_points");
  launch_reset_slots(ag, persist, make_pool(ctx), s);
  // phase A: sample agents (skipped when paused / using a link), targets, characters
  launch_sample_points(ctx->nav, *o, n, ag.position, ag.flags, LMB_AGENT_PAUSED | LMB_AGENT_USING_ANIMATION_LINK,
                       0, ctx->d_agent_point.p, ctx->d_agent_node.p, s);
  launch_sample_points(ctx->nav, *o, n, ag.target, ag.flags, LMB_AGENT_PAUSED | LMB_AGENT_USING_ANIMATION_LINK,
                       LMB_AGENT_HAS_TARGET, ctx->d_target_point.p, ctx->d_target_node.p, s);
  launch_sample_points(ctx->nav, *o, ctx->n_chars, ctx->d_char_position.p, nullptr, 0, 0, ctx->d_char_point.p,
                       ctx->d_char_node.p, s);
  T.kernel_launches += (n ? 3 : 0) + (ctx->n_chars ? 1 : 0);
  nvtxRangePop();
  cudaEventRecord(ctx->ev[3], s);

  // phase B: decision
  RepathQueue queue{};
  queue.count = ctx->d_queue_count.p;
  queue.agent = ctx->d_q_agent.p;
  queue.slot = ctx->d_q_slot.p;
  queue.start_node = ctx->d_q_start_node.p;
  queue.end_node = ctx->d_q_end_node.p;
  queue.start_point = ctx->d_q_start_point.p;
  queue.end_point = ctx->d_q_end_point.p;
  FollowScratch scratch{};
  scratch.follow_start = ctx->d_follow_start.p;
  scratch.follow_end = ctx->d_follow_end.p;
  scratch.presult = ctx->d_presult.p;
  scratch.query_of_agent = ctx->d_query_of_agent.p;
  scratch.waypoint_kind = ctx->d_waypoint_kind.p;
  scratch.waypoint_link = ctx->d_waypoint_link.p;
  scratch.waypoint_points = ctx->d_waypoint_points.p;
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_queue_count.p, 0, sizeof(uint32_t), s));
  nvtxRangePushA("lmb:phaseB repath_decide");
  launch_repath_decide(ctx->nav, ag, persist, make_pool(ctx), queue, scratch, s);
  nvtxRangePop();
  T.kernel_launches += n ? 1 : 0;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_small, ctx->d_queue_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  cudaEventRecord(ctx->ev[4], s);
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  const uint32_t n_repath = ctx->h_small[0];
  T.n_repath = n_repath;

  // phase B: A*
  AStarQueries q = make_queries(ctx, n_repath);
  q.owner = ctx->d_q_agent.p;
  q.slot = ctx->d_q_slot.p;
  if (ctx->has_overrides) {
    q.override_begin = ctx->d_override_begin.p;
    q.override_type = ctx->d_override_type.p;
    q.override_cost = ctx->d_override_cost.p;
  }
  q.owner_flags = ctx->d_flags.p;
  if (ctx->has_permitted) {
    q.permitted_begin = ctx->d_permitted_begin.p;
    q.permitted_kind = ctx->d_permitted_kind.p;
  }
  if (n_repath) {
    NvtxRange r("lmb:phaseB astar");
    st = run_astar(ctx, q);
    if (st != LMB_OK) return st;
  }
  cudaEventRecord(ctx->ev[5], s);

  // phase C
  nvtxRangePushA("lmb:phaseC follow");
  // from a few thousand agents on, k_follow takes them longest corridor first (three small launches)
  if (n >= 8192 || (n && (ctx->cfg.flags & LMB_CONFIG_ORDER_ALWAYS))) {
    CUDA_TRY(ctx, ctx->d_follow_order.reserve(n));
    CUDA_TRY(ctx, ctx->d_follow_key.reserve(n));
    CUDA_TRY(ctx, ctx->d_follow_hist.reserve(2 * FOLLOW_BUCKETS));
    launch_follow_order(ag, make_pool(ctx), q, scratch, ctx->d_follow_key.p, ctx->d_follow_hist.p, ctx->d_follow_order.p, s);
    scratch.order = ctx->d_follow_order.p;
    T.kernel_launches += 3;
  }
  launch_follow(ctx->nav, ag, persist, make_pool(ctx), q, scratch, s);
  nvtxRangePop();
  T.kernel_launches += n ? 1 : 0;
  CUDA_TRY(ctx, cudaGetLastError());
  cudaEventRecord(ctx->ev[6], s);

  // phase D, first half: this rank's avoidance records into the halo buffer
  ctx->halo_total = n_total;
  ctx->halo_first = first;
  (void)dt;
  if (!(ctx->cfg.flags & LMB_CONFIG_NO_AVOIDANCE)) {
    NvtxRange r("lmb:phaseD records");
    st = lmb_avoidance_build_records(ctx, o, n_total, first);
    if (st != LMB_OK) return st;
  }
  cudaEventRecord(ctx->ev[10], s);
  if (sync_at_end) CUDA_TRY(ctx, cudaStreamSynchronize(s));  // the caller may now all-gather the halo buffer
  ctx->step_open = true;
  return LMB_OK;
}

extern "C" {

int32_t lmb_halo_buffer(lmb_ctx* ctx, void** device_ptr, uint64_t* record_bytes, uint64_t* n_records) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->step_open) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_halo_buffer: no step in flight");
  if (device_ptr) *device_ptr = ctx->d_records.p;
  if (record_bytes) *record_bytes = sizeof(AvoidRecord);
  if (n_records) *n_records = ctx->halo_total;
  return LMB_OK;
}

int32_t lmb_step_end(lmb_ctx* ctx, const lmb_options* o, float dt) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->step_open) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step_end without lmb_step_begin");
  if (!o) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step_end: null options");
  ctx->step_open = false;
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  const uint32_t n = ctx->n_agents;
  lmb_step_timings& T = ctx->timings;
  AgentArrays ag = lmb_make_agent_arrays(ctx);
  AgentPersist persist = lmb_make_persist(ctx);
  int32_t st;
  cudaEventRecord(ctx->ev[9], s);
  // phase D, second half: grid over ALL records, ORCA for this rank's agents
  const bool avoid = !(ctx->cfg.flags & LMB_CONFIG_NO_AVOIDANCE) && n > 0;
  ctx->h_small[8] = 0;
  for (int attempt = 0;; attempt++) {
    if (avoid) {
      NvtxRange r("lmb:phaseD avoidance");
      st = lmb_avoidance_solve(ctx, o, dt, ctx->halo_total, ctx->halo_first);
      if (st != LMB_OK) return st;
    }
    cudaEventRecord(ctx->ev[7], s);
    launch_gather_outputs(ag, persist, ctx->d_out_velocity.p, ctx->d_out_state.p, ctx->d_out_link_id.p,
                          ctx->d_out_link_start.p, ctx->d_out_link_end.p, s);
    T.kernel_launches += n ? 1 : 0;
    CUDA_TRY(ctx, cudaGetLastError());
    cudaEventRecord(ctx->ev[8], s);
    CUDA_TRY(ctx, cudaStreamSynchronize(s));  // the step's one synchronisation behind the A* phase
    if (!avoid || ctx->h_small[8] == 0) break;
    // some agent's avoidance scratch was too small: nothing was committed; grow for good and redo phase D
    if (attempt >= 7) return ctx->fail(LMB_ERR_CAPACITY, "avoidance scratch kept overflowing");
    for (int k = 0; k < 5; k++) ctx->avoid_caps[k] *= 2;
  }
  cudaEventElapsedTime(&T.sample_ms, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&T.repath_decide_ms, ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&T.astar_ms, ctx->ev[4], ctx->ev[5]);
  cudaEventElapsedTime(&T.follow_ms, ctx->ev[5], ctx->ev[6]);
  float a1 = 0.0f, a2 = 0.0f;
  cudaEventElapsedTime(&a1, ctx->ev[6], ctx->ev[10]);  // record building
  cudaEventElapsedTime(&a2, ctx->ev[9], ctx->ev[7]);   // grid + solve
  T.avoidance_ms = a1 + a2;
  // device time of the step excluding whatever the caller did between begin and end (halo exchange)
  float t1 = 0.0f, t2 = 0.0f;
  cudaEventElapsedTime(&t1, ctx->ev[2], ctx->ev[10]);
  cudaEventElapsedTime(&t2, ctx->ev[9], ctx->ev[8]);
  T.total_ms = t1 + t2;
  ctx->last_n_results = n;
  return LMB_OK;
}

int32_t lmb_step_resident(lmb_ctx* ctx, const lmb_options* o, float dt) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  // (no caller stands between the halves here: phase D is enqueued behind phase C without a host sync)
  int32_t st = lmb_step_begin_impl(ctx, o, dt, ctx->n_agents, 0, /*sync_at_end=*/false);
  if (st != LMB_OK) return st;
  return lmb_step_end(ctx, o, dt);
}

int32_t lmb_agents_integrate(lmb_ctx* ctx, float delta_time) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (ctx->last_n_results != ctx->n_agents)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_integrate: no step has run for the uploaded agents");
  cudaSetDevice(ctx->device);
  launch_integrate(ctx->n_agents, delta_time, ctx->d_out_velocity.p, ctx->d_position.p, ctx->d_velocity.p,
                   ctx->d_flags.p, ctx->stream);
  CUDA_TRY(ctx, cudaGetLastError());
  return LMB_OK;
}

int32_t lmb_agents_positions(lmb_ctx* ctx, float* out_position, uint32_t capacity) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (ctx->n_agents > capacity) return ctx->fail(LMB_ERR_CAPACITY, "lmb_agents_positions: %u agents", ctx->n_agents);
  if (ctx->n_agents == 0) return LMB_OK;
  if (!out_position) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_positions: null output");
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaMemcpyAsync(out_position, ctx->d_position.p, 3 * (size_t)ctx->n_agents * sizeof(float),
                                cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return LMB_OK;
}

int32_t lmb_agents_download(lmb_ctx* ctx, const lmb_agents_out* out) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agents_download: null out");
  const uint32_t n = ctx->n_agents;
  if (n == 0) return LMB_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  cudaEventRecord(ctx->ev[0], s);
  if (out->desired_velocity)
    CUDA_TRY(ctx, cudaMemcpyAsync(out->desired_velocity, ctx->d_out_velocity.p, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out->state)
    CUDA_TRY(ctx, cudaMemcpyAsync(out->state, ctx->d_out_state.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (out->reached_link_id)
    CUDA_TRY(ctx, cudaMemcpyAsync(out->reached_link_id, ctx->d_out_link_id.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (out->reached_link_start)
    CUDA_TRY(ctx, cudaMemcpyAsync(out->reached_link_start, ctx->d_out_link_start.p, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out->reached_link_end)
    CUDA_TRY(ctx, cudaMemcpyAsync(out->reached_link_end, ctx->d_out_link_end.p, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
  cudaEventRecord(ctx->ev[1], s);
  CUDA_TRY(ctx, cudaStreamSynchronize(s));
  cudaEventElapsedTime(&ctx->timings.d2h_ms, ctx->ev[0], ctx->ev[1]);
  return LMB_OK;
}

int32_t lmb_step(lmb_ctx* ctx, const lmb_agents_in* agents, const lmb_characters_in* characters,
                 const lmb_options* options, float delta_time, const lmb_agents_out* out) {
  int32_t st = lmb_agents_upload(ctx, agents, characters);
  if (st != LMB_OK) return st;
  st = lmb_step_resident(ctx, options, delta_time);
  if (st != LMB_OK) return st;
  float h2d = ctx->timings.h2d_ms;
  st = lmb_agents_download(ctx, out);
  ctx->timings.h2d_ms = h2d;
  return st;
}

int32_t lmb_pathing_results(lmb_ctx* ctx, lmb_pathing_result* out, uint32_t capacity, uint32_t* out_count) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_count) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_pathing_results: null out_count");
  *out_count = ctx->timings.n_repath;
  if (ctx->timings.n_repath > capacity) return LMB_ERR_CAPACITY;
  if (ctx->timings.n_repath == 0) return LMB_OK;
  if (!out) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_pathing_results: null out");
  cudaSetDevice(ctx->device);
  const uint32_t n = ctx->last_n_results;
  std::vector<uint32_t> pr(n);
  CUDA_TRY(ctx, cudaMemcpy(pr.data(), ctx->d_presult.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint32_t k = 0;
  for (uint32_t i = 0; i < n && k < capacity; i++)
    if (pr[i] & 0x80000000u) out[k++] = {i, (pr[i] >> 30) & 1u, pr[i] & 0x3fffffffu};
  *out_count = k;
  return LMB_OK;
}

int32_t lmb_agent_path(lmb_ctx* ctx, uint32_t slot, uint32_t* out_nodes, uint32_t* out_portals,
                       uint32_t capacity, uint32_t* out_len) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_len || slot >= ctx->cfg.max_agents)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_path: bad slot / null out_len");
  cudaSetDevice(ctx->device);
  uint32_t off = 0, len = 0;
  CUDA_TRY(ctx, cudaMemcpy(&len, ctx->d_slot_path_len.p + slot, 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(ctx, cudaMemcpy(&off, ctx->d_slot_path_off.p + slot, 4, cudaMemcpyDeviceToHost));
  *out_len = len;
  if (!len) return LMB_OK;
  if (len > capacity) return LMB_ERR_CAPACITY;
  std::vector<uint2> tmp(len);
  CUDA_TRY(ctx, cudaMemcpy(tmp.data(), ctx->d_pool[ctx->pool_cur].p + off, len * sizeof(uint2), cudaMemcpyDeviceToHost));
  if (tmp[0].x == LMB_NONE) {  // invalidated by the last lmb_navdata_update, not yet seen by a step
    *out_len = 0;
    return LMB_OK;
  }
  for (uint32_t k = 0; k < len; k++) {
    out_nodes[k] = tmp[k].x;
    out_portals[k] = tmp[k].y;
  }
  return LMB_OK;
}

int32_t lmb_agent_path_endpoints(lmb_ctx* ctx, uint32_t slot, float* out_start_point, float* out_end_point) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_start_point || !out_end_point || slot >= ctx->cfg.max_agents)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_path_endpoints: bad slot / null output");
  cudaSetDevice(ctx->device);
  float ends[6];
  CUDA_TRY(ctx, cudaMemcpy(ends, ctx->d_path_endpoints.p + 6 * (size_t)slot, sizeof(ends), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 3; k++) {
    out_start_point[k] = ends[k];
    out_end_point[k] = ends[3 + k];
  }
  return LMB_OK;
}

int32_t lmb_agent_waypoints(lmb_ctx* ctx, uint32_t capacity, uint32_t* out_kind, float* out_points,
                            uint32_t* out_link, uint32_t* out_count) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_count) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_waypoints: null out_count");
  const uint32_t n = ctx->n_agents;
  *out_count = n;
  if (n > capacity) return LMB_ERR_CAPACITY;
  if (n == 0) return LMB_OK;
  if (!out_kind || !out_points || !out_link)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_waypoints: null output");
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaMemcpy(out_kind, ctx->d_waypoint_kind.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(ctx, cudaMemcpy(out_link, ctx->d_waypoint_link.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(ctx, cudaMemcpy(out_points, ctx->d_waypoint_points.p, 6 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return LMB_OK;
}

int32_t lmb_agent_avoidance_constraints(lmb_ctx* ctx, uint32_t agent_index, uint32_t capacity, float* out_lines,
                                        uint32_t* out_n_original, uint32_t* out_n_fallback, uint32_t* out_ran) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_n_original || !out_n_fallback || !out_ran)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_avoidance_constraints: null output");
  if (agent_index >= ctx->n_agents)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_avoidance_constraints: agent %u of %u", agent_index,
                     ctx->n_agents);
  *out_n_original = *out_n_fallback = *out_ran = 0;
  if (ctx->h_dbg_row_of.empty() || ctx->h_dbg_row_of[agent_index] == LMB_NONE || !ctx->dbg_lines_per_row)
    return LMB_OK;  // the agent did not ask for its avoidance data (or no step ran since the upload)
  const uint32_t row = ctx->h_dbg_row_of[agent_index];
  cudaSetDevice(ctx->device);
  uint32_t h[4];
  CUDA_TRY(ctx, cudaMemcpy(h, ctx->d_dbg_header.p + 4 * (size_t)row, sizeof(h), cudaMemcpyDeviceToHost));
  *out_ran = h[0] ? 1u + h[3] : 0u;
  *out_n_original = h[1];
  *out_n_fallback = h[2];
  const uint32_t total = h[1] + h[2];
  if (total > capacity) return LMB_ERR_CAPACITY;
  if (total) {
    if (!out_lines) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_agent_avoidance_constraints: null out_lines");
    CUDA_TRY(ctx, cudaMemcpy(out_lines, ctx->d_dbg_lines.p + (size_t)row * ctx->dbg_lines_per_row,
                             (size_t)total * 4 * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return LMB_OK;
}

int32_t lmb_get_step_timings(lmb_ctx* ctx, lmb_step_timings* out) {
  if (!ctx || !out) return LMB_ERR_INVALID_ARGUMENT;
  *out = ctx->timings;
  return LMB_OK;
}

int32_t lmb_get_query_timings(lmb_ctx* ctx, lmb_step_timings* out) {
  if (!ctx || !out) return LMB_ERR_INVALID_ARGUMENT;
  *out = ctx->query_timings;
  return LMB_OK;
}

}  // extern "C"
