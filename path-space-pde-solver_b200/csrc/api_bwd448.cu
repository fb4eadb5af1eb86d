// api_bwd448.cu -- detached-backward kernel instantiation for 448 threads per CTA.
#include "api_common.h"

int pspde_launch_bwd_448(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  return launch_rollout<448, true, 1>(pl, p, stream);
}
