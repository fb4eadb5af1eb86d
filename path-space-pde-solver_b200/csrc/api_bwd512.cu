// api_bwd512.cu -- template instantiations for T = 512 threads per CTA.
#include "api_common.h"

int pspde_launch_bwd_512(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  if (pl.NB == 2) return launch_rollout<512, true, 2>(pl, p, stream);
  if (pl.NB == 3) return launch_rollout<512, true, 3>(pl, p, stream);
  return fail(-13, "internal: no backward kernel for T=%d NB=%d", pl.T, pl.NB);
}
