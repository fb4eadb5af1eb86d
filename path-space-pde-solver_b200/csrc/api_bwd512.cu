// api_bwd512.cu -- detached-backward kernel instantiation for 512 threads per CTA.
#include "api_common.h"

int pspde_launch_bwd_512(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  return launch_rollout<512, true, 1>(pl, p, stream);
}
