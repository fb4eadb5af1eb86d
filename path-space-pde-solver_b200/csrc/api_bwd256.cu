// api_bwd256.cu -- detached-backward kernel instantiation for 256 threads per CTA.
#include "api_common.h"

int pspde_launch_bwd_256(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  return launch_rollout<256, true, 1>(pl, p, stream);
}
