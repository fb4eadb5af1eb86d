// api_bwd256.cu -- detached-backward kernel instantiation for 256 threads per CTA.
#include "api_common.h"
#include "grad_kernels.cuh"

int pspde_launch_bwd_256(const Plan& pl, const pspde::RolloutParams& p, void* stream) {
  return launch_rollout<256, true, 1>(pl, p, stream);
}

int pspde_launch_grad_256(const Plan& pl, const pspde::RolloutParams& p, int grid, int n_items, void* stream) {
  if (!p.th_tbl) return fail(-13, "could not allocate the weight-image index table");
  auto kern = pspde::grad_kernel<kP, 256>;
  if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute(%zu B smem) failed", pl.smem_bytes);
  PSPDE_LAUNCH(kern, grid, 256, pl.smem_bytes, stream, p, n_items);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "gradient kernel launch failed: %s", e);
  return 0;
}
