// rollout_tc_kernels.cuh -- forward rollout on the 5th-generation tensor cores (tcgen05 + tensor memory).
//
// Same computation as rollout_kernel<BWD = false> (reference hot loop solver.py:440-494) for the shape class
//   DenseNet (function_space.py:116-140) or MySequential (:177-195) with two hidden layers of at most 32 (31) units,
//   'inner' time input, diagonal problem functors (LLGC / LQGC with off_diag = 0, DoubleWell_multidim),
// which covers the BASELINE configs C1-inner, C2, C3 (detached) and C5.  A tile is 128 trajectories = the 128 lanes of tensor
// memory; every trajectory is owned by kTcTPP = 4 threads (one per quarter of the state columns) for the whole rollout:
//
//   tensor memory (512 columns x 128 lanes, all allocated):
//     A operands, hi and lo TF32 halves: a0 = [X | t | 1 | 0] (s0 columns), h1 (32), h2 (32)       2 (s0 + 64) columns
//     accumulator D = [pre1 (32) | pre2 (32) | Z (np3)]                                              64 + np3 columns
//   shared memory: the weights as six K-major B tiles (hi and lo of B0, B1, B2), staged once per CTA
//     B0 = [W0 | W1[a0 rows] | W2[a0 rows]]   (s0 x (64 + np3))    the dense-concat layers all read a0: ONE MMA group
//     B1 = [W1[h1 rows] | W2[h1 rows]]        (32 x (32 + np3))
//     B2 =  W2[h2 rows]                       (32 x np3)
//   registers: the step's Brownian increments (own columns), the per-path partial sums of Y, Z_sum, g; the state X lives in
//     the a0 operand columns of tensor memory (hi + lo is exact) and is read back for the Euler-Maruyama update.
//
// Per step: G0 = a0.B0 -> h1 = relu(pre1)^2 -> G1 += h1.B1 -> h2 = relu(pre2)^2 -> G2 += h2.B2 -> Z; the
// Euler-Maruyama update then runs on the registers of the thread that owns the column and writes the next a0
// straight back to tensor memory (tcgen05.st): activations never take a shared-memory layout.  Each product is
// FP32-equivalent through the 3-pass hi/lo split of tc_sm100.cuh.  One thread (kTcIssuer) issues the MMAs; the hand-offs are
// mbarriers (operands ready: every thread arrives; group done: tcgen05.commit).
#pragma once
#if !defined(PSPDE_EMULATE)
#include <mutex>
#include "rollout_kernels.cuh"
#include "tc_sm100.cuh"

namespace pspde {

constexpr int kTcP = 128;        // trajectories per tile = tensor-memory lanes
// Threads per trajectory (column parts).  FOUR, i.e. 16 warps per CTA at 128 registers (one CTA per SM: 213 KB of weights).
// History: while the state X lived in registers, four threads per trajectory spilled ~100 values per thread (local memory is an
// L2 round trip here: L1 is what shared memory leaves, ~14 KB) and THREE (170 registers) was faster.  Since X is read back from
// its tensor-memory operand columns the kernel spills ~40 - 50 values either way, and the fourth warp per scheduler hides more
// of the tensor-memory / mbarrier latency than the spills cost: row-keeping forward 4.60 -> 4.39 ms at C2, C3 step 70.0 ->
// 66.1 ms, C5 step 372 -> 368 ms (the plain forward is a wash: 105 -> 108 ms at C5).
constexpr int kTcTPP = 4;
constexpr int kTcThreads = kTcP * kTcTPP;
constexpr int kTcIssuer = (kTcThreads / 32 - 1) * 32; // the thread that issues the MMAs: a lane of the last warp, whose column part is
                                   // the short one (the state columns do not divide evenly), so the issue work does not delay the slowest warp
constexpr int kTcMaxG = 12;      // float4 column groups per thread; the kernel is instantiated for NG <= this

struct TcGeom {
  int s0, hp, np3, n0, n1, n2, ng, dense;   // ng = column groups per thread (max over the parts)
  int c_a0h, c_a0l, c_h1h, c_h1l, c_h2h, c_h2l, c_d;                  // tensor-memory columns
  uint32_t o_b0h, o_b0l, o_b1h, o_b1l, o_b2h, o_b2l, o_prob, o_exch, o_red, total;   // shared-memory byte offsets
};

// false if the network / problem is outside the shape class of this kernel
inline bool tc_geom(const NetGeom& g, int d, TcGeom& t) {
  if (g.L != 3 || g.time_mode != TIME_FIRST || g.dims[3] != d) return false;
  const bool dense = g.kind == NET_DENSENET;
  if (g.seg_len[1] > 32 || g.seg_len[2] > 32) return false;     // MySequential: hidden width + its bias column <= 32
  t.s0 = (g.seg_len[0] + 7) & ~7;
  t.hp = 32;
  t.np3 = (t.s0 + 15) & ~15;        // >= s0: the SDE step may read Z for every own column group
  // DenseNet: every layer reads a0, so group g also produces the a0 / h1 part of the later layers (accumulated on);
  // MySequential: one layer per group
  t.n0 = dense ? 2 * t.hp + t.np3 : t.hp; t.n1 = dense ? t.hp + t.np3 : t.hp; t.n2 = t.np3;
  t.dense = dense ? 1 : 0;
  if (t.n0 > 256) return false;
  t.ng = (t.s0 / 4 + kTcTPP - 1) / kTcTPP;
  if (t.ng > kTcMaxG) return false;
  t.c_a0h = 0; t.c_a0l = t.s0; t.c_h1h = 2 * t.s0; t.c_h1l = t.c_h1h + t.hp; t.c_h2h = t.c_h1l + t.hp;
  t.c_h2l = t.c_h2h + t.hp; t.c_d = t.c_h2l + t.hp;
  if (t.c_d + 2 * t.hp + t.np3 > 512) return false;
  uint32_t o = 0;
  t.o_b0h = o; o += tc::b_tile_bytes(t.s0, t.n0);
  t.o_b0l = o; o += tc::b_tile_bytes(t.s0, t.n0);
  t.o_b1h = o; o += tc::b_tile_bytes(t.hp, t.n1);
  t.o_b1l = o; o += tc::b_tile_bytes(t.hp, t.n1);
  t.o_b2h = o; o += tc::b_tile_bytes(t.hp, t.n2);
  t.o_b2l = o; o += tc::b_tile_bytes(t.hp, t.n2);
  t.o_prob = o; o += 7u * (uint32_t)t.s0 * 4u;
  t.o_exch = o; o += (uint32_t)kTcP * (uint32_t)kTcTPP * 4u * 4u;
  t.o_red = o;  o += 4u * 8u;
  t.total = o;
  return t.total <= 227u * 1024u;
}

// weights of one B tile: row k <-> activation column row0 + k (NetGeom numbering), column n <-> (layer, column) by range
__device__ __forceinline__ void tc_stage_tile(const NetGeom& g, const float* __restrict__ th, uint8_t* hi, uint8_t* lo,
                                              int Kp, int Np, int row0, int seg_len, int l_first, int hp, int tid, int nthr) {
  for (int q = tid; q < Kp * Np; q += nthr) {
    const int k = q / Np, n = q - k * Np;
    // columns [0, hp) belong to layer l_first unless it is the output layer; then hp more for the next hidden layer ...
    int l = l_first, nn = n;
    while (l < g.L - 1 && nn >= hp) { nn -= hp; ++l; }
    float w = 0.f;
    if (k < seg_len) {
      const int idx = theta_index(g, l, (g.kind == NET_DENSENET ? row0 : 0) + k, nn);   // row relative to the layer's first input column
      if (idx >= 0) w = __ldg(th + idx);
    }
    float h, r;
    tc::tf32_split_rn(w, h, r);
    *reinterpret_cast<float*>(hi + tc::b_tile_offset(n, k, Np)) = h;
    *reinterpret_cast<float*>(lo + tc::b_tile_offset(n, k, Np)) = r;
  }
}

// The six B tiles exactly as they sit in shared memory ([o_b0h, o_prob): hi / lo of B0, B1, B2), built ONCE per launch in
// global memory; every CTA of the rollout then fetches the image with one bulk copy (cp.async.bulk -> mbarrier) instead of
// evaluating theta_index for its own 53 k elements.
static __global__ void tc_pack_weights_kernel(const NetGeom g, const float* __restrict__ theta, const TcGeom tg, uint8_t* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  tc_stage_tile(g, theta, out + tg.o_b0h, out + tg.o_b0l, tg.s0, tg.n0, 0, g.seg_len[0], 0, tg.hp, tid, nthr);
  tc_stage_tile(g, theta, out + tg.o_b1h, out + tg.o_b1l, tg.hp, tg.n1, g.seg_off[1], g.seg_len[1], 1, tg.hp, tid, nthr);
  tc_stage_tile(g, theta, out + tg.o_b2h, out + tg.o_b2l, tg.hp, tg.n2, g.seg_off[2], g.seg_len[2], 2, tg.hp, tid, nthr);
}

// streaming store of one checkpoint element (written once, read once by another kernel)
__device__ __forceinline__ void st_ckpt(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// One trajectory = one tensor-memory lane, owned by kTcTPP threads (one per part of the state columns).  Thread kTcIssuer
// additionally issues the MMAs: after the last arrival on an "operands ready" barrier it launches the group and
// commits it to the matching "group done" barrier; everybody (the issuer included) then waits for that one.
// CKPT = true (detached backward, first half): the same rollout, but instead of the per-path outputs it writes the
// operand rows of the gradient accumulation for every step -- the network input a0 = [X_n | t_n | 1], the hidden
// activations h1, h2 and the cotangent on Z, zeta = wY (sqrt(dt) xi + [!adaptive] Z dt) + wZ Z dt -- to the
// per-wave checkpoint buffer (RolloutParams::ckpt, column-major: one row of 128 paths per column) that the gradient
// kernels consume.  Rows with zero cotangents (padding,
// trajectories dropped by the host because their D was non-finite) are written as zeros: inert in the gradient.
// DIAG = true: additionally the u_L2 diagnostic of solver.py:491-494 from the per-step device tables (include/pspde.h).
// PHILOX is a template parameter on purpose: with the noise source a run-time flag the 4 NG 64-bit addresses of the injected
// increments are step-loop invariants, get hoisted, and push the Philox instantiation over the register limit.
template <int NG, bool CKPT, bool DIAG, bool PHILOX>
__global__ void __launch_bounds__(kTcThreads, 1) rollout_tc_fwd_kernel(const RolloutParams prm, const TcGeom tg) {
  extern __shared__ float4 smem4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  __shared__ uint64_t bars[7];          // [0..2] operands of group g ready (all threads arrive), [3..5] group g done (commit),
                                        // [6] the weight image has landed (bulk copy)
  __shared__ uint32_t tmem_base_s;
  const NetGeom& g = prm.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = prm.d, N = prm.N;
  const float dt = prm.dt, sq = sqrtf(prm.dt);
  double* sRed = reinterpret_cast<double*>(smem + tg.o_red);
  float* sProb = reinterpret_cast<float*>(smem + tg.o_prob);
  float* sExch = reinterpret_cast<float*>(smem + tg.o_exch);

  // ---- one-time setup: tensor memory, barriers, weights (hi / lo B tiles), problem vectors
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { tc::mbar_init(&bars[i], kTcThreads); tc::mbar_init(&bars[3 + i], 1); }
    tc::mbar_init(&bars[6], 1);
    tc::mbar_fence_init();
    if (prm.wpack) {      // weights: ONE bulk copy (TMA engine) of the packed image, completion on bars[6]
      tc::mbar_arrive_expect_tx(&bars[6], tg.o_prob);
      tc::bulk_load(smem + tg.o_b0h, prm.wpack, tg.o_prob, &bars[6]);
    }
  }
  if (tid < 4) sRed[tid] = 0.0;
  if (!prm.wpack) {       // no packed image (allocation failed): every CTA builds its tiles from theta
    tc_stage_tile(g, prm.theta, smem + tg.o_b0h, smem + tg.o_b0l, tg.s0, tg.n0, 0, g.seg_len[0], 0, tg.hp, tid, kTcThreads);
    tc_stage_tile(g, prm.theta, smem + tg.o_b1h, smem + tg.o_b1l, tg.hp, tg.n1, g.seg_off[1], g.seg_len[1], 1, tg.hp, tid, kTcThreads);
    tc_stage_tile(g, prm.theta, smem + tg.o_b2h, smem + tg.o_b2l, tg.hp, tg.n2, g.seg_off[2], g.seg_len[2], 2, tg.hp, tid, kTcThreads);
  }
  for (int q = tid; q < 7 * tg.s0; q += kTcThreads) {
    const int v = q / tg.s0, j = q - v * tg.s0;
    sProb[q] = j < d ? __ldg(prm.prob + v * d + j) : 0.f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (prm.wpack) tc::mbar_wait(&bars[6], 0u);
  const uint32_t tbase = tmem_base_s;

  // ---- MMA groups (thread kTcIssuer)
  const uint32_t sb = tc::smem_u32(smem);
  const uint32_t dcol = tbase + (uint32_t)tg.c_d;
  auto issue = [&](int grp, uint32_t parity) {
    tc::mbar_wait(&bars[grp], parity);
    tc::fence_after_sync();
    if (grp == 0)
      tc::mma_3xtf32(dcol, tbase + tg.c_a0h, tbase + tg.c_a0l, sb + tg.o_b0h, sb + tg.o_b0l, tg.n0, tg.s0 / 8,
                     tc::idesc_tf32(128, tg.n0), false, (uint32_t)tg.n0 * 16u, 128u);
    else if (grp == 1)
      tc::mma_3xtf32(dcol + tg.hp, tbase + tg.c_h1h, tbase + tg.c_h1l, sb + tg.o_b1h, sb + tg.o_b1l, tg.n1, tg.hp / 8,
                     tc::idesc_tf32(128, tg.n1), tg.dense != 0, (uint32_t)tg.n1 * 16u, 128u);
    else
      tc::mma_3xtf32(dcol + 2 * tg.hp, tbase + tg.c_h2h, tbase + tg.c_h2l, sb + tg.o_b2h, sb + tg.o_b2l, tg.n2, tg.hp / 8,
                     tc::idesc_tf32(128, tg.n2), tg.dense != 0, (uint32_t)tg.n2 * 16u, 128u);
    tc::mma_commit(&bars[3 + grp]);
  };

  // ---- thread = (trajectory, column part)
  const int qtr = warp & 3, part = warp >> 2;
  const int p = 32 * qtr + lane;
  const uint32_t lane_addr = ((uint32_t)(32 * qtr)) << 16;
  const uint32_t tA0h = tbase + lane_addr + tg.c_a0h, tA0l = tbase + lane_addr + tg.c_a0l;
  const uint32_t tD = tbase + lane_addr + tg.c_d;
  const int G = tg.s0 / 4, gbase = G / kTcTPP, grem = G % kTcTPP;
  const int g_lo = part * gbase + (part < grem ? part : grem);
  const int ng = gbase + (part < grem ? 1 : 0);
  // hidden columns per thread in the activation epilogues: the 8 column quads of a hidden segment go 3 / 3 / 2 to the parts
  constexpr int HQ = (8 + kTcTPP - 1) / kTcTPP, HC = 4 * HQ;
  const int hq0 = HQ * part, hnq = (8 - hq0) < HQ ? (8 - hq0) : HQ;
  const float *a_d = sProb, *b_d = sProb + tg.s0, *p_d = sProb + 2 * tg.s0, *r_d = sProb + 3 * tg.s0,
              *al = sProb + 4 * tg.s0, *kap = sProb + 5 * tg.s0, *eta = sProb + 6 * tg.s0;
  const bool adaptive = prm.adaptive != 0, dw = prm.problem_id == PROBLEM_DW;
  uint32_t ph = 0;

  for (int tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
    const int k = (prm.tile0 + tile) * kTcP + p;
    const bool in = k < prm.K_local;
    float wy = 0.f, wz = 0.f;
    if (CKPT && in) {
      if (prm.ckpt_unit) wy = 1.0f;      // forward pass that keeps its operand rows: cotangents are applied by the gradient kernel
      else { wy = prm.wY ? __ldg(prm.wY + k) : 0.f; wz = prm.wZ ? __ldg(prm.wZ + k) : 0.f; }
    }
    const bool live = CKPT && (wy != 0.f || wz != 0.f);
    const bool keep = CKPT && tile < prm.ckpt_tiles;       // a forward pass may keep the rows of its first tiles only
    // checkpoint rows of this path: column col of step n at ck[(n * ckpt_cols + col) * 128] (a warp stores 128 contiguous bytes)
    float* ck = CKPT ? prm.ckpt + (size_t)tile * N * prm.ckpt_cols * kTcP + p : nullptr;
    const unsigned kglob = (unsigned)(prm.k_offset + k);
    // ---- tile init (solver.py:365-376) and the first a0
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      if (gi < ng) {
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * (g_lo + gi) + i;
          float x = 0.f;
          if (j < d) { if (in) x = prm.x0_per_path ? __ldg(prm.x0 + (size_t)k * d + j) : __ldg(prm.x0 + j); }
          else if (j == d) x = prm.t_index ? (float)__ldg(prm.t_index) * prm.dt_net : 0.f;    // network time of step 0
          else if (j == d + 1) x = 1.0f;
          tc::tf32_split(x, hi[i], lo[i]);
        }
        tc::tmem_st4(tA0h + 4 * (g_lo + gi), hi);
        tc::tmem_st4(tA0l + 4 * (g_lo + gi), lo);
      }
    }
    tc::wait_st();
    tc::fence_before_sync();
    tc::mbar_arrive(&bars[0]);
    if (tid == kTcIssuer) issue(0, ph);
    float yp = (part == 0 && prm.y0) ? __ldg(prm.y0) : 0.f, zsp = 0.f, gp = 0.f, fip = 0.f, ulp = 0.f;
    PhaseTimer pt_;          // debug: [0,2,4] wait for MMA group 0/1/2, [1,3] hidden epilogues, [5] SDE step, [6] noise
    pt_.start(prm.prof, tid == 32 ? 0 : 1);      // an ordinary thread (kTcIssuer also issues the MMAs)

    for (int n = 0; n < N; ++n) {
      const bool last = (n == N - 1);
      // ---- Brownian increments of this step for the own columns.  They do not depend on the network, so they are
      //      generated (branch-free, NG independent Philox chains) while the tensor core works on G0.
      // The draw is split in three parts placed in front of the three waits for the tensor core (G0 is the longest).
      float E[NG][4];
      constexpr int NA = (NG + 1) / 2, NB = NA + (NG - NA + 1) / 2;
      auto draw = [&](int g0, int g1) {
        if (PHILOX) {
          // The chains are independent, and the scheduler would interleave all of them (about 14 live registers each).  Two
          // at a time hide the multiply latency just as well: `dep` (always 0, but only at run time -- a normal is never that
          // NaN pattern) makes every chain wait for the one two places before it.
          unsigned dep[2] = {0u, 0u};
#pragma unroll
          for (int gi = 0; gi < NG; ++gi) {
            if (gi >= g0 && gi < g1) {
              const float4 e4 = philox_normal4(kglob, (unsigned)n, (unsigned)(g_lo + gi), prm.offset + dep[gi & 1], prm.seed);
              E[gi][0] = e4.x; E[gi][1] = e4.y; E[gi][2] = e4.z; E[gi][3] = e4.w;
              dep[gi & 1] = (__float_as_uint(e4.w) == 0x7fffffffu) ? 1u : 0u;
            }
          }
        } else {
#pragma unroll
          for (int gi = 0; gi < NG; ++gi)
            if (gi >= g0 && gi < g1) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int j = 4 * (g_lo + gi) + i;
                E[gi][i] = (in && gi < ng && j < d) ? __ldg(prm.xi + (long long)k * prm.xs_k + (long long)j * prm.xs_j + (long long)n * prm.xs_n) : 0.f;
              }
            }
        }
      };
      draw(0, NA);
      pt_.mark(6);
      // ---- hidden layers: pre -> h = relu(.)^2 -> A operand (hi, lo)
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) {
        tc::mbar_wait(&bars[3 + hl], ph);
        tc::fence_after_sync();
        pt_.mark(2 * hl);
        float v[HC], hi[HC], lo[HC];
#pragma unroll
        for (int u = 0; u < HQ; ++u) {
          float v4[4] = {0.f, 0.f, 0.f, 0.f};
          if (u < hnq) tc::tmem_ld4(tD + hl * tg.hp + 4 * (hq0 + u), v4);
          v[4 * u] = v4[0]; v[4 * u + 1] = v4[1]; v[4 * u + 2] = v4[2]; v[4 * u + 3] = v4[3];
        }
        tc::wait_ld();
        if (tg.dense) {
#pragma unroll
          for (int i = 0; i < HC; ++i) { const float s = fmaxf(v[i], 0.f); tc::tf32_split(s * s, hi[i], lo[i]); }
        } else {                           // MySequential: tanh, and the constant-1 column that carries the next bias
          const int one_col = g.dims[1 + hl];
#pragma unroll
          for (int i = 0; i < HC; ++i) {
            const float h = (4 * hq0 + i == one_col) ? 1.0f : tanhf(v[i]);
            tc::tf32_split(h, hi[i], lo[i]);
          }
        }
        const uint32_t th = tbase + lane_addr + (hl ? tg.c_h2h : tg.c_h1h) + 4 * hq0;
#pragma unroll
        for (int u = 0; u < HQ; ++u) {
          if (u < hnq) {
            const float h4[4] = {hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]};
            const float l4[4] = {lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]};
            tc::tmem_st4(th + 4 * u, h4);
            tc::tmem_st4(th + tg.hp + 4 * u, l4);
          }
        }
        if (CKPT && keep) {                  // h = hi + lo exactly
          float* o = ck + (size_t)(n * prm.ckpt_cols + tg.s0 + tg.hp * hl + 4 * hq0) * kTcP;
#pragma unroll
          for (int u = 0; u < HC; ++u)
            if (u < 4 * hnq) st_ckpt(o + u * kTcP, live ? hi[u] + lo[u] : 0.f);
        }
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&bars[1 + hl]);
        if (tid == kTcIssuer) issue(1 + hl, ph);
        if (hl == 0) draw(NA, NB); else draw(NB, NG);
        pt_.mark(2 * hl + 1);
      }
      // ---- Z is complete: Euler-Maruyama step on the own columns (solver.py:471-486)
      tc::mbar_wait(&bars[5], ph);
      tc::fence_after_sync();
      pt_.mark(4);
      // Branch-free per element: the problem vectors are zero on the [t | 1 | pad] columns (so x stays put there) and Z
      // is exactly zero on them (zero weight columns); only the time column is patched afterwards.
      float zz = 0.f, zxi = 0.f, ff = 0.f, ul = 0.f, gg = 0.f;
      // network time of the next step: (n + 1) dt, or the caller's grid (importance sampling, Solver.Z_n :360-362)
      const float t_next = (prm.t_index && !last) ? (float)__ldg(prm.t_index + n + 1) * prm.dt_net : (float)(n + 1) * dt;
      const float cm = adaptive ? -1.0f : 0.f;
      // checkpoint rows of this step: first own column of a0 / of zeta (one warp-wide store = 128 contiguous bytes)
      float* ckx = CKPT ? ck + (size_t)(n * prm.ckpt_cols + 4 * g_lo) * kTcP : nullptr;
      const int zoff = (tg.s0 + 2 * tg.hp) * kTcP;
#pragma unroll
      for (int c0 = 0; c0 < NG; c0 += 2) {          // 2 column groups (8 columns) per tensor-memory access
        const bool full = (c0 + 1 < NG) && (c0 + 1 < ng);     // warp-uniform
        // The state X_n is NOT kept in registers across the step: it is read back from the a0 operand columns of tensor
        // memory (hi + lo == x exactly, tc::tf32_split) together with Z.
        float Z[8], H[8], Lo[8];
        if (full) {
          tc::tmem_ld8(tD + 2 * tg.hp + 4 * (g_lo + c0), Z);
          tc::tmem_ld8(tA0h + 4 * (g_lo + c0), H);
          tc::tmem_ld8(tA0l + 4 * (g_lo + c0), Lo);
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float z4[4] = {0.f, 0.f, 0.f, 0.f}, h4[4] = {0.f, 0.f, 0.f, 0.f}, l4[4] = {0.f, 0.f, 0.f, 0.f};
            if (c0 + u < NG && c0 + u < ng) {
              tc::tmem_ld4(tD + 2 * tg.hp + 4 * (g_lo + c0 + u), z4);
              tc::tmem_ld4(tA0h + 4 * (g_lo + c0 + u), h4);
              tc::tmem_ld4(tA0l + 4 * (g_lo + c0 + u), l4);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { Z[4 * u + i] = z4[i]; H[4 * u + i] = h4[i]; Lo[4 * u + i] = l4[i]; }
          }
        }
        tc::wait_ld();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (c0 + u < NG) {
            const int gi = c0 + u;
            const int j0 = 4 * (g_lo + gi);
            if (gi < ng) {
              const float4 A4 = ld4(a_d + j0), B4 = ld4(b_d + j0), P4 = ld4(p_d + j0), K4 = ld4(kap + j0);
              const float av[4] = {A4.x, A4.y, A4.z, A4.w}, bv[4] = {B4.x, B4.y, B4.z, B4.w};
              const float pv[4] = {P4.x, P4.y, P4.z, P4.w}, kv[4] = {K4.x, K4.y, K4.z, K4.w};
              if (CKPT && keep) {            // operand rows of this step: a0 = X_n (own columns) and, unless the gradient
                // kernel regenerates it, zeta; group gi of this thread sits 4 gi columns (an immediate offset) behind its first one
#pragma unroll
                for (int i = 0; i < 4; ++i) st_ckpt(ckx + (4 * gi + i) * kTcP, live ? H[4 * u + i] + Lo[4 * u + i] : 0.f);
                if (prm.ckpt_zeta) {
                  const float kA = adaptive ? 0.f : dt;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float z = Z[4 * u + i], ee = E[gi][i];
                    const float ze = (j0 + i < d) ? wy * (sq * ee + kA * z) + wz * (dt * z) : 0.f;
                    st_ckpt(ckx + zoff + (4 * gi + i) * kTcP, live ? ze : 0.f);
                  }
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float z = Z[4 * u + i], x = H[4 * u + i] + Lo[4 * u + i], ee = E[gi][i];
                zz = fmaf(z, z, zz);
                zxi = fmaf(z, ee, zxi);
                const float drift = fmaf(av[i], x, -(4.0f * kv[i] * (x * (x * x - 1.0f))));   // OU: kappa = 0; double well: a = 0
                float xn = x + (drift + bv[i] * (cm * z)) * dt + (bv[i] * ee) * sq;
                ff = fmaf(pv[i] * xn, xn, ff);
                if (DIAG && j0 + i < d) {                                                  // (-Z - u*(X_{n+1}, t_n))^2, solver.py:492-493
                  const int j = j0 + i;
                  float us;
                  if (prm.u_mode == 1) us = __ldg(prm.u_tab + (size_t)(2 * n) * d + j) + __ldg(prm.u_tab + (size_t)(2 * n + 1) * d + j) * xn;
                  else {
                    const float xc = fminf(fmaxf(xn, -prm.u_xb), prm.u_xb - 2.0f * prm.u_dx);
                    int cell = (int)floorf((xc + prm.u_xb) / prm.u_dx);
                    cell = cell < 0 ? 0 : (cell >= prm.u_nx1 ? prm.u_nx1 - 1 : cell);
                    if (prm.k_offset + k == prm.u_quirk) { cell -= 2; if (cell < 0) cell += prm.u_nx1; }   // `i[-1] -= 2`, problems.py:279
                    us = __ldg(prm.u_tab + ((size_t)(2 * n) + (j < prm.u_d1 ? 0 : 1)) * prm.u_nx1 + cell);
                  }
                  const float du = -z - us;
                  ul = fmaf(du, du, ul);
                }
                xn = (j0 + i == d) ? t_next : xn;                                          // the time column
                if (last) {                                                                // terminal cost g(X_N), X_N itself
                  if (j0 + i < d) {
                    gg += al[j0 + i] * xn + r_d[j0 + i] * xn * xn + eta[j0 + i] * (xn - 1.0f) * (xn - 1.0f);
                    if (prm.X_N && in) prm.X_N[(size_t)k * d + j0 + i] = xn;
                  }
                }
                tc::tf32_split(xn, H[4 * u + i], Lo[4 * u + i]);
              }
            }
          }
        }
        if (!last) {
          if (full) {
            tc::tmem_st8(tA0h + 4 * (g_lo + c0), H);
            tc::tmem_st8(tA0l + 4 * (g_lo + c0), Lo);
          } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (c0 + u < NG && c0 + u < ng) {
                const float h4[4] = {H[4 * u], H[4 * u + 1], H[4 * u + 2], H[4 * u + 3]};
                const float l4[4] = {Lo[4 * u], Lo[4 * u + 1], Lo[4 * u + 2], Lo[4 * u + 3]};
                tc::tmem_st4(tA0h + 4 * (g_lo + c0 + u), h4);
                tc::tmem_st4(tA0l + 4 * (g_lo + c0 + u), l4);
              }
            }
          }
        }
      }
      const float run = 0.5f * zz + ff;
      yp += (run + (adaptive ? -zz : 0.f)) * dt + zxi * sq;
      zsp += run * dt;
      fip += ff * dt;
      if (DIAG) ulp += ul * dt;
      if (last) gp = gg;
      else {
        tc::wait_st();
        tc::fence_before_sync();
        tc::mbar_arrive(&bars[0]);
      }
      ph ^= 1u;
      if (!last && tid == kTcIssuer) issue(0, ph);
      pt_.mark(5);
    }

    // ---- tile epilogue: combine the column parts of every trajectory, outputs, statistics
    if (part != 0) {
      float* e = sExch + 4 * (kTcP * (part - 1) + p);
      e[0] = yp; e[1] = zsp; e[2] = gp; e[3] = fip;
      if (DIAG) sExch[4 * kTcP * (kTcTPP - 1) + kTcP * (part - 1) + p] = ulp;
    }
    __syncthreads();
    if (part == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      if (in) {
        float Y = yp, ZS = zsp, Gv = gp, FI = fip, UL = ulp;
#pragma unroll
        for (int r = 1; r < kTcTPP; ++r) {
          const float* e = sExch + 4 * (kTcP * (r - 1) + p);
          Y += e[0]; ZS += e[1]; Gv += e[2]; FI += e[3];
          if (DIAG) UL += sExch[4 * kTcP * (kTcTPP - 1) + kTcP * (r - 1) + p];
        }
        if (DIAG && prm.uL2) prm.uL2[k] = UL;
        const double D = (double)Y - (double)Gv;
        const bool keep = path_kept(D, ZS, prm.d_abs_max);
        if (prm.Y_N) prm.Y_N[k] = keep ? Y : dropped_mark(Y);
        if (prm.gX) prm.gX[k] = Gv;
        if (prm.Zsum) prm.Zsum[k] = ZS;
        if (prm.Fint) prm.Fint[k] = FI;
        if (keep) { s0 = D; s1 = D * D; s2 = (double)ZS + (double)Gv; }
        else s3 = 1.0;
      }
      s0 = warp_sum_d(s0); s1 = warp_sum_d(s1); s2 = warp_sum_d(s2); s3 = warp_sum_d(s3);
      if (lane == 0) { atomicAdd(sRed + 0, s0); atomicAdd(sRed + 1, s1); atomicAdd(sRed + 2, s2); atomicAdd(sRed + 3, s3); }
    }
    __syncthreads();      // sExch is rewritten by the next tile
  }

  tc::fence_before_sync();
  __syncthreads();
  if (tid < 4 && prm.stats_partial) prm.stats_partial[blockIdx.x * 4 + tid] = sRed[tid];
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// Scratch for the packed weight image, one per (device, stream) that ever launched a rollout (at most 16, 256 KB each, kept for
// the life of the process): stream order alone then keeps a pack kernel from overwriting an image a rollout is still loading.
// nullptr (table full or allocation failure) = the rollout stages its tiles itself.
inline void* tc_wpack_buffer(size_t bytes, cudaStream_t stream) {
  struct Slot { int dev; cudaStream_t stream; void* buf; size_t bytes; };
  static Slot slots[16];
  static int n_slots = 0;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_slots; ++i)
    if (slots[i].dev == dev && slots[i].stream == stream) {
      if (slots[i].bytes >= bytes) return slots[i].buf;
      cudaFree(slots[i].buf);                       // a larger network on this stream: grow (cudaFree synchronises)
      if (cudaMalloc(&slots[i].buf, bytes) != cudaSuccess) { (void)cudaGetLastError(); slots[i].buf = nullptr; slots[i].bytes = 0; return nullptr; }
      slots[i].bytes = bytes;
      return slots[i].buf;
    }
  if (n_slots == 16) return nullptr;
  void* buf = nullptr;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  slots[n_slots++] = Slot{dev, stream, buf, bytes};
  return buf;
}

// NG instantiations: the smallest one that holds tg.ng column groups per thread
template <int NG, bool CKPT, bool DIAG, bool PHILOX>
inline cudaError_t tc_launch_k(const RolloutParams& p, const TcGeom& tg, int grid, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(rollout_tc_fwd_kernel<NG, CKPT, DIAG, PHILOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tg.total);
  if (e != cudaSuccess) return e;
  RolloutParams q = p;
  q.wpack = nullptr;
  if (void* wp = tc_wpack_buffer(tg.o_prob, stream)) {               // scratch for the packed weight image (per device and stream)
    tc_pack_weights_kernel<<<64, 256, 0, stream>>>(p.g, p.theta, tg, reinterpret_cast<uint8_t*>(wp));
    g_launches++;
    q.wpack = reinterpret_cast<const uint8_t*>(wp);
  }
  rollout_tc_fwd_kernel<NG, CKPT, DIAG, PHILOX><<<grid, kTcThreads, tg.total, stream>>>(q, tg);
  return cudaGetLastError();
}
template <int NG, bool CKPT, bool DIAG>
inline cudaError_t tc_launch_one(const RolloutParams& p, const TcGeom& tg, int grid, cudaStream_t stream) {
  return p.noise_mode == NOISE_PHILOX ? tc_launch_k<NG, CKPT, DIAG, true>(p, tg, grid, stream)
                                      : tc_launch_k<NG, CKPT, DIAG, false>(p, tg, grid, stream);
}
template <bool CKPT, bool DIAG = false>
inline cudaError_t tc_launch_t(const RolloutParams& p, const TcGeom& tg, int grid, cudaStream_t stream) {
  if (tg.ng <= 2) return tc_launch_one<2, CKPT, DIAG>(p, tg, grid, stream);
  if (tg.ng <= 4) return tc_launch_one<4, CKPT, DIAG>(p, tg, grid, stream);      // (C3: d = 50)
  if (tg.ng <= 7) return tc_launch_one<7, CKPT, DIAG>(p, tg, grid, stream);      // (C2 / C5: d = 100)
  return tc_launch_one<kTcMaxG, CKPT, DIAG>(p, tg, grid, stream);
}
inline cudaError_t tc_launch(const RolloutParams& p, const TcGeom& tg, int grid, cudaStream_t stream) {
  return p.u_mode != 0 ? tc_launch_t<false, true>(p, tg, grid, stream) : tc_launch_t<false, false>(p, tg, grid, stream);
}
// forward pass that also leaves the operand rows of ALL its tiles in prm.ckpt (unit cotangents, RolloutParams::ckpt_unit)
inline cudaError_t tc_launch_fwd_ckpt(const RolloutParams& p, const TcGeom& tg, int grid, cudaStream_t stream) {
  return p.u_mode != 0 ? tc_launch_t<true, true>(p, tg, grid, stream) : tc_launch_t<true, false>(p, tg, grid, stream);
}
// columns per (tile slot, step) of the checkpoint buffer: a0 (s0) | h1 (hp) | h2 (hp) | zeta (s0)
inline int tc_ckpt_cols(const TcGeom& tg) { return 2 * tg.s0 + 2 * tg.hp; }

}  // namespace pspde
#endif  // !PSPDE_EMULATE
