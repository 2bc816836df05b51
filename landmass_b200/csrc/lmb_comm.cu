// landmass_b200 — the library-owned halo exchange of the multi-GPU step.
//
// Agents shard over ranks by contiguous ranges, navigation data is replicated, and the ONE exchange
// of a step is the all-gather of the 32-byte avoidance records between phase C and phase D
// (avoidance.rs:84-130 needs every agent's sampled point / velocity / radius).  lmb_step_begin /
// lmb_halo_buffer / lmb_step_end leave that exchange to the caller; here the library does it itself:
// ncclAllGather enqueued on the context's stream between the two halves, no host synchronisation
// around it.  NCCL is opened at run time (dlopen of libnccl.so.2 — the copy already loaded by the
// host process if there is one), so the library has no link-time dependency on it and single-GPU
// callers never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "lmb_context.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

void load_nccl(NcclApi& api);

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { load_nccl(api); });
  return api;
}

}  // namespace

namespace {
void load_nccl(NcclApi& api) {
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.error = std::string("cannot open libnccl.so.2: ") + dlerror();
    return;
  }
  auto sym = [&](const char* n) -> void* {
    void* p = dlsym(api.handle, n);
    if (!p && api.error.empty()) api.error = std::string("libnccl: missing symbol ") + n;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!api.error.empty()) api.handle = nullptr;
}

}  // namespace

int32_t lmb_step_begin_impl(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t n_total, uint32_t first,
                            bool sync_at_end);  // lmb_api.cu

void lmb_comm_release(lmb_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  NcclApi& api = nccl_api();
  if (api.handle) api.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->comm_ranks = 0;
}

extern "C" {

int32_t lmb_comm_unique_id(lmb_ctx* ctx, uint8_t out_id[128]) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!out_id) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_comm_unique_id: null output");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi& api = nccl_api();
  if (!api.handle) return ctx->fail(LMB_ERR_UNSUPPORTED, "%s", api.error.c_str());
  ncclUniqueId id;
  const ncclResult_t r = api.GetUniqueId(&id);
  if (r != ncclSuccess) return ctx->fail(LMB_ERR_CUDA, "ncclGetUniqueId: %s", api.GetErrorString(r));
  memcpy(out_id, &id, sizeof(id));
  return LMB_OK;
}

int32_t lmb_comm_init(lmb_ctx* ctx, const uint8_t id_bytes[128], uint32_t n_ranks, uint32_t rank) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!id_bytes || n_ranks == 0 || rank >= n_ranks)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_comm_init: bad arguments (rank %u of %u)", rank, n_ranks);
  NcclApi& api = nccl_api();
  if (!api.handle) return ctx->fail(LMB_ERR_UNSUPPORTED, "%s", api.error.c_str());
  lmb_comm_release(ctx);
  cudaSetDevice(ctx->device);
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api.CommInitRank(&comm, (int)n_ranks, id, (int)rank);
  if (r != ncclSuccess) return ctx->fail(LMB_ERR_CUDA, "ncclCommInitRank: %s", api.GetErrorString(r));
  ctx->comm = comm;
  ctx->comm_ranks = n_ranks;
  ctx->comm_rank = rank;
  return LMB_OK;
}

int32_t lmb_step_sharded(lmb_ctx* ctx, const lmb_options* o, float dt, uint32_t agents_per_rank) {
  if (!ctx) return LMB_ERR_INVALID_ARGUMENT;
  if (!ctx->comm) return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step_sharded: lmb_comm_init has not been called");
  if (ctx->n_agents > agents_per_rank)
    return ctx->fail(LMB_ERR_INVALID_ARGUMENT, "lmb_step_sharded: %u agents on this rank exceed agents_per_rank %u",
                     ctx->n_agents, agents_per_rank);
  if ((uint64_t)agents_per_rank * ctx->comm_ranks > 0xFFFFFFF0ull)
    return ctx->fail(LMB_ERR_CAPACITY, "lmb_step_sharded: more than 2^32 records");
  NcclApi& api = nccl_api();
  const uint32_t n_total = agents_per_rank * ctx->comm_ranks, first = agents_per_rank * ctx->comm_rank;
  int32_t st = lmb_step_begin_impl(ctx, o, dt, n_total, first, /*sync_at_end=*/false);
  if (st != LMB_OK) return st;
  cudaStream_t s = ctx->stream;
  cudaEventRecord(ctx->ev_ag[0], s);
  if (!(ctx->cfg.flags & LMB_CONFIG_NO_AVOIDANCE)) {
    AvoidRecord* mine = ctx->d_records.p + first;
    // the unused tail of this rank's share: valid = 0
    if (ctx->n_agents < agents_per_rank)
      CUDA_TRY(ctx, cudaMemsetAsync(mine + ctx->n_agents, 0,
                                    (size_t)(agents_per_rank - ctx->n_agents) * sizeof(AvoidRecord), s));
    const ncclResult_t r = api.AllGather(mine, ctx->d_records.p, (size_t)agents_per_rank * sizeof(AvoidRecord), ncclChar,
                                         static_cast<ncclComm_t>(ctx->comm), s);
    if (r != ncclSuccess) {
      ctx->step_open = false;
      return ctx->fail(LMB_ERR_CUDA, "ncclAllGather: %s", api.GetErrorString(r));
    }
  }
  cudaEventRecord(ctx->ev_ag[1], s);
  st = lmb_step_end(ctx, o, dt);
  if (st != LMB_OK) return st;
  cudaEventElapsedTime(&ctx->timings.allgather_ms, ctx->ev_ag[0], ctx->ev_ag[1]);
  ctx->timings.total_ms += ctx->timings.allgather_ms;  // step_end's total leaves out what lies between the halves
  return LMB_OK;
}

}  // extern "C"
