// landmass_b200 — phase D host orchestration: records -> uniform grid -> per-agent
// obstacle flood + ORCA (k_avoid.cu).  Split in two so that a multi-GPU caller can
// all-gather the per-agent avoidance records between the halves.
#include <algorithm>
#include <cmath>

#include "lmb_context.h"

namespace lmb {
size_t avoid_scratch_bytes_per_agent(const AvoidScratch& s);
void launch_avoid_grid(uint32_t n_records, const AvoidRecord* records, AvoidGrid& grid, uint32_t* max_radius_bits,
                       cudaStream_t stream);
void launch_stage_moves(const AgentArrays& ag, const AgentPersist& persist, float* staged, cudaStream_t stream);
void launch_commit_moves(const AgentArrays& ag, const AgentPersist& persist, const float* staged,
                         const uint32_t* overflow, cudaStream_t stream);
void launch_avoidance_chunk(const DevNav& nav, const AgentArrays& ag, const lmb_options& o, float delta_time,
                            uint32_t first_agent_global, uint32_t chunk_begin, uint32_t chunk_n,
                            uint32_t n_records, const AvoidRecord* records, uint32_t n_characters,
                            const AvoidRecord* char_records, const AvoidGrid& grid, const AvoidScratch& scratch,
                            const uint32_t* max_radius_bits, float* staged_moves, const AvoidDebug& dbg,
                            const AvoidOrder& ord, cudaStream_t stream);
void launch_avoid_order(const AvoidGrid& grid, uint32_t n_records, uint32_t first, uint32_t n, const AvoidOrder& ord,
                        cudaStream_t stream);
}  // namespace lmb

AgentArrays lmb_make_agent_arrays(lmb_ctx* ctx);
AgentPersist lmb_make_persist(lmb_ctx* ctx);

// Builds this rank's avoidance records into ctx->d_records[first .. first + n_agents).
int32_t lmb_avoidance_build_records(lmb_ctx* ctx, const lmb_options* o, uint32_t n_total, uint32_t first) {
  cudaStream_t s = ctx->stream;
  CUDA_TRY(ctx, ctx->d_records.reserve(n_total ? n_total : 1));
  AgentArrays ag = lmb_make_agent_arrays(ctx);
  launch_build_records(ag, lmb_make_persist(ctx), *o, ctx->d_records.p + first, s);
  if (ctx->n_chars) {
    CUDA_TRY(ctx, ctx->d_char_records.reserve(ctx->n_chars));
    launch_character_records(ctx->n_chars, ctx->d_char_velocity.p, ctx->d_char_radius.p, ctx->d_char_point.p,
                             ctx->d_char_node.p, ctx->d_char_records.p, s);
  }
  ctx->timings.kernel_launches += (ag.n ? 1 : 0) + (ctx->n_chars ? 1 : 0);
  CUDA_TRY(ctx, cudaGetLastError());
  return LMB_OK;
}

// Enqueues avoidance for this rank's agents against all n_total records: grid, per-agent solve into a staging
// buffer, commit.  No host synchronisation: if some agent's scratch overflows, the device-side flag keeps the
// commit kernel from publishing anything and the caller (lmb_step_end) sees the flag at its own final sync,
// grows the capacities and calls this again — the capacities stick, so a steady state never retries.
int32_t lmb_avoidance_solve(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first) {
  cudaStream_t s = ctx->stream;
  const uint32_t n = ctx->n_agents;
  if (n == 0) return LMB_OK;
  AgentArrays ag = lmb_make_agent_arrays(ctx);
  AgentPersist persist = lmb_make_persist(ctx);

  // uniform grid over the navmesh's world xy bounds
  AvoidGrid grid{};
  float cell = std::max(o->neighbourhood * 0.5f, 1e-3f);
  float ex = std::max(ctx->world_max[0] - ctx->world_min[0], 1e-3f) + 2.0f * cell;
  float ey = std::max(ctx->world_max[1] - ctx->world_min[1], 1e-3f) + 2.0f * cell;
  const double max_cells = 8.0e6;
  if ((double)(ex / cell) * (double)(ey / cell) > max_cells) cell = (float)std::sqrt((double)ex * ey / max_cells);
  grid.ox = ctx->world_min[0] - cell;
  grid.oy = ctx->world_min[1] - cell;
  grid.inv_cell = 1.0f / cell;
  grid.nx = (uint32_t)std::ceil(ex / cell) + 1;
  grid.ny = (uint32_t)std::ceil(ey / cell) + 1;
  const size_t cells = (size_t)grid.nx * grid.ny;
  CUDA_TRY(ctx, ctx->d_cell_count.reserve(cells + 2));
  CUDA_TRY(ctx, ctx->d_cell_start2.reserve(cells + 2));
  grid.cell_count = ctx->d_cell_count.p;
  grid.cell_start = ctx->d_cell_start2.p;
  CUDA_TRY(ctx, ctx->d_sorted_records.reserve(n_total ? n_total : 1));
  grid.sorted_rec = ctx->d_sorted_records.p;
  CUDA_TRY(ctx, ctx->d_scan_tiles.reserve(exclusive_scan_tmp_words((uint32_t)cells + 1)));
  grid.scan_tmp = ctx->d_scan_tiles.p;
  uint32_t* max_radius_bits = ctx->d_avoid_overflow.p + 1;
  launch_avoid_grid(n_total, ctx->d_records.p, grid, max_radius_bits, s);
  ctx->timings.kernel_launches += 4;
  CUDA_TRY(ctx, cudaGetLastError());

  CUDA_TRY(ctx, ctx->d_out_velocity.reserve(3 * (size_t)n));
  float* staged = ctx->d_out_velocity.p;  // reused as the staging buffer; gather overwrites it later

  {
    AvoidScratch sc{};
    sc.max_neighbours = ctx->avoid_caps[0];
    sc.max_explore = ctx->avoid_caps[1];
    sc.max_border_edges = ctx->avoid_caps[2];
    sc.max_cones = ctx->avoid_caps[3];
    sc.max_lines = ctx->avoid_caps[4];
    sc.per_agent_bytes = avoid_scratch_bytes_per_agent(sc);
    const uint32_t chunk = std::min<uint32_t>(n, 131072u);
    sc.agent_stride = (chunk + 31u) & ~31u;
    CUDA_TRY(ctx, ctx->d_avoid_scratch.reserve(sc.per_agent_bytes * (size_t)sc.agent_stride));
    sc.base = ctx->d_avoid_scratch.p;
    {
      // The flood's explored table must be EMPTY (0xFFFFFFFF) when an agent starts; floods clean up behind
      // themselves, so the whole scratch is filled once per layout (capacities, stride or buffer changed).
      const uint64_t sig = ((uint64_t)sc.agent_stride << 32) ^ ((uint64_t)sc.max_explore << 20) ^ ((uint64_t)sc.max_lines << 12) ^
                           ((uint64_t)sc.max_cones << 6) ^ sc.max_border_edges ^ ((uint64_t)sc.max_neighbours << 40) ^
                           (uint64_t)(uintptr_t)sc.base;
      if (sig != ctx->avoid_scratch_sig) {
        CUDA_TRY(ctx, cudaMemsetAsync(sc.base, 0xFF, sc.per_agent_bytes * (size_t)sc.agent_stride, s));
        ctx->avoid_scratch_sig = sig;
      }
    }
    sc.overflow = ctx->d_avoid_overflow.p;
    CUDA_TRY(ctx, cudaMemsetAsync(sc.overflow, 0, sizeof(uint32_t), s));
    // debug-avoidance export: one row of 2 x max_lines per asking agent, "did not run" until written
    AvoidDebug dbg{};
    ctx->dbg_lines_per_row = 0;
    if (ctx->dbg_rows) {
      ctx->dbg_lines_per_row = 2 * sc.max_lines;
      CUDA_TRY(ctx, ctx->d_dbg_header.reserve(4 * (size_t)ctx->dbg_rows));
      CUDA_TRY(ctx, ctx->d_dbg_lines.reserve((size_t)ctx->dbg_rows * ctx->dbg_lines_per_row));
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_dbg_header.p, 0, 4 * (size_t)ctx->dbg_rows * sizeof(uint32_t), s));
      dbg.row_of = ctx->d_dbg_row_of.p;
      dbg.header = ctx->d_dbg_header.p;
      dbg.lines = ctx->d_dbg_lines.p;
      dbg.lines_per_row = ctx->dbg_lines_per_row;
    }
    // processing order (AvoidOrder): cell order from a few thousand agents on
    AvoidOrder ord{};
    if (n >= 8192 || (ctx->cfg.flags & LMB_CONFIG_ORDER_ALWAYS)) {
      if (n_total == n && first == 0) {
        ord.mode = 1;
      } else {
        ord.mode = 2;
        CUDA_TRY(ctx, ctx->d_avoid_flags.reserve((size_t)n_total + 1));
        CUDA_TRY(ctx, ctx->d_avoid_pos.reserve((size_t)n_total + 1));
        CUDA_TRY(ctx, ctx->d_avoid_order.reserve(n));
        CUDA_TRY(ctx, ctx->d_avoid_scan.reserve(exclusive_scan_tmp_words(n_total + 1)));
        ord.flags = ctx->d_avoid_flags.p;
        ord.pos = ctx->d_avoid_pos.p;
        ord.order = ctx->d_avoid_order.p;
        ord.scan_tmp = ctx->d_avoid_scan.p;
        launch_avoid_order(grid, n_total, first, n, ord, s);
        ctx->timings.kernel_launches += 4;
      }
    }
    launch_stage_moves(ag, persist, staged, s);
    for (uint32_t b = 0; b < n; b += chunk) {
      launch_avoidance_chunk(ctx->nav, ag, *o, dt, first, b, std::min(chunk, n - b), n_total, ctx->d_records.p,
                             ctx->n_chars, ctx->d_char_records.p, grid, sc, max_radius_bits, staged, dbg, ord, s);
      ctx->timings.kernel_launches += 1;
    }
    launch_commit_moves(ag, persist, staged, sc.overflow, s);
    ctx->timings.kernel_launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_small + 8, sc.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  }
  return LMB_OK;
}

