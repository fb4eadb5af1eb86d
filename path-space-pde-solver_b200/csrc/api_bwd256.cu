// api_bwd256.cu -- template instantiations for T = 256 threads per CTA.
#include "api_common.h"

int pspde_launch_bwd_256(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  if (pl.NB == 1) return launch_rollout<256, true, 1>(pl, p, stream);
  if (pl.NB == 2) return launch_rollout<256, true, 2>(pl, p, stream);
  if (pl.NB == 3) return launch_rollout<256, true, 3>(pl, p, stream);
  if (pl.NB == 8) return launch_rollout<256, true, 8>(pl, p, stream);
  return fail(-13, "internal: no backward kernel for T=%d NB=%d", pl.T, pl.NB);
}
