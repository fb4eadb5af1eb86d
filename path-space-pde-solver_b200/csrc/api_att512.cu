// api_att512.cu -- attached-mode kernel instantiation for 512 threads per CTA.
#include "api_common.h"

int pspde_launch_att_512(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  if (!p.th_tbl) return fail(-13, "could not allocate the weight-image index table");
  auto kern = rollout_attached_kernel<kP, 512, 1>;
  if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed");
  PSPDE_LAUNCH(kern, pl.grid, 512, pl.smem_bytes, stream, p);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
  return 0;
}
