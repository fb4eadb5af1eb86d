// api_att256.cu -- attached-mode kernel instantiation for 256 threads per CTA.
#include "api_common.h"

int pspde_launch_att_256(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  if (!p.th_tbl) return fail(-13, "could not allocate the weight-image index table");
  if (pl.ctas_per_sm == 2) {     // two CTAs per SM (16 warps): <= 128 registers per thread
    auto kern2 = rollout_attached_kernel<kP, 256, 1, 2>;
    if (pspde_set_smem(kern2, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed");
    PSPDE_LAUNCH(kern2, pl.grid, 256, pl.smem_bytes, stream, p);
    g_launches++;
    if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
    return 0;
  }
  auto kern = rollout_attached_kernel<kP, 256, 1>;
  if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed");
  PSPDE_LAUNCH(kern, pl.grid, 256, pl.smem_bytes, stream, p);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
  return 0;
}
