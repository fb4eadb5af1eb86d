// api_att256.cu -- template instantiations for T = 256 threads per CTA.
#include "api_common.h"

int pspde_launch_att_256(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
#define PSPDE_ATT(NBB)                                                                       \
  {                                                                                         \
    auto kern = rollout_attached_kernel<kP, 256, NBB>;                                      \
    if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed"); \
    PSPDE_LAUNCH(kern, pl.grid, 256, pl.smem_bytes, stream, p);                              \
  }
  if (pl.NB == 1) PSPDE_ATT(1)
  else if (pl.NB == 2) PSPDE_ATT(2)
  else if (pl.NB == 3) PSPDE_ATT(3)
  else if (pl.NB == 8) PSPDE_ATT(8)
  else return fail(-13, "internal: no attached kernel for T=%d NB=%d", pl.T, pl.NB);
#undef PSPDE_ATT
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
  return 0;
}
