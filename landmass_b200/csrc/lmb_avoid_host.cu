// landmass_b200 — phase D host orchestration (placeholder until k_avoid.cu lands).
#include "lmb_context.h"

extern "C" int32_t lmb_avoidance_phase(lmb_ctx* ctx, const lmb_options* o, float dt) {
  (void)o;
  (void)dt;
  return ctx->fail(LMB_ERR_UNSUPPORTED, "avoidance kernels not built yet");
}
