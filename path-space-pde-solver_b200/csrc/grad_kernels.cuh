// grad_kernels.cuh -- second half of the checkpointed detached backward (reference: loss.backward(), solver.py:221,
// for detach_forward=True).
//
// With a detached forward process the gradient is a plain sum over (trajectory, step) samples,
//     dLoss/dtheta = sum_{k,n} J_theta Z(t_n, X_{k,n})' zeta_{k,n}          (oracle/manual.py::grad_mode_a),
// and every sample needs only its operand row [a0 | h1 | h2 | zeta]: the hidden cotangents are
//     delta_2 = (zeta W2[h2 rows]') * act'(h2),  delta_1 = (zeta W2[h1 rows]' + delta_2 W1[h1 rows]') * act'(h1)
// and dW_l += act_l' delta_l.  The tensor-core forward kernel (rollout_tc_fwd_kernel<.., CKPT = true>) regenerates the
// trajectories of one WAVE of tiles (at most one 128-path tile per SM) and leaves those rows in the checkpoint
// buffer; this kernel streams them back (L2 / HBM, 1 088 B per sample at the C2 shape; column-major rows of 128 paths), P = 64 samples at a time,
// and accumulates the weight gradient in registers exactly like rollout_kernel<BWD = true> does (same routines:
// net_backward_hidden, bw_accum, bw_flush).  The buffer is per wave, so its size does not depend on K.
// The samples are independent: work items (slot, step, half tile) are dealt round-robin to the CTAs.
#pragma once
#include "rollout_kernels.cuh"

namespace pspde {

constexpr int kCkP = 128;       // paths per checkpoint tile (= the tensor-core kernel's tile)

template <int P, int T>
__global__ void __launch_bounds__(T, 1) grad_kernel(const RolloutParams prm, const int n_items) {
  PSPDE_DYN_SMEM(smem4);
  float* smem = reinterpret_cast<float*>(smem4);
  const NetGeom& g = prm.g;
  const SmemLayout sl = smem_layout(g, P, true, false);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = T / 32;
  constexpr int HALVES = kCkP / P;
  float* sAct = smem + sl.act;
  float* sZe = smem + sl.xi;

  for (int q = sl.act + tid; q < sl.total; q += T) smem[q] = 0.f;
  __syncthreads();
  stage_weights(g, prm.theta, smem + sl.w, tid, T);
  f32x2 acc[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) acc[q] = f2_zero();
  const BwSlot slot = bw_slot(g, P, tid, T);
  float* gp = prm.grad_partial + (size_t)blockIdx.x * prm.n_img_total;

  // checkpoint columns -> tiles: segment s of the activation row, then zeta (column-major rows of 128 paths)
  const int s0 = prm.ckpt_s0;
  const int seg_src[3] = {0, s0, s0 + 32};
  const int ze_src = s0 + 64, ze_n = ceil4(prm.d);
  const float* ck = prm.ckpt;
  int since_flush = 0;
  __syncthreads();

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int ts = item / HALVES, half = item - ts * HALVES;
    const float* src = ck + (size_t)ts * prm.ckpt_cols * kCkP + half * P;
    for (int s = 0; s < g.L; ++s) {
      float* dst = sAct + g.seg_off[s];
      for (int q = tid; q < g.seg_len[s] * P; q += T) {
        const int c = q / P, p = q - c * P;
        dst[p * g.lda + c] = __ldg(src + (size_t)(seg_src[s] + c) * kCkP + p);
      }
    }
    for (int q = tid; q < ze_n * P; q += T) {
      const int c = q / P, p = q - c * P;
      sZe[p * g.ldz + c] = c < prm.ckpt_s0 ? __ldg(src + (size_t)(ze_src + c) * kCkP + p) : 0.f;
    }
    __syncthreads();
    net_backward_hidden<P>(prm, sl, smem, warp, lane, NW);
    bw_accum<P>(acc, g, sl, smem, slot);
    if (++since_flush == 128) { bw_flush(acc, g, slot, gp, lane); since_flush = 0; }   // bounds the fp32 accumulation length
    __syncthreads();
  }
  bw_flush(acc, g, slot, gp, lane);
}

}  // namespace pspde
