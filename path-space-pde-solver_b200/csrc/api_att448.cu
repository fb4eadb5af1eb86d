// api_att448.cu -- attached-mode kernel instantiation for 448 threads per CTA.
#include "api_common.h"

int pspde_launch_att_448(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  auto kern = rollout_attached_kernel<kP, 448, 1>;
  if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed");
  PSPDE_LAUNCH(kern, pl.grid, 448, pl.smem_bytes, stream, p);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
  return 0;
}
