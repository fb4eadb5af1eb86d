// grad_tc2_kernels.cuh -- gradient accumulation from the checkpoint rows with EVERYTHING on the tensor cores: the weight
// gradient and the hidden cotangents (reference: loss.backward(), solver.py:221, for detach_forward=True).
//
// Contract (same as grad_tc_kernel): a streaming kernel over (trajectory, step) samples,
//     dtheta = sum_samples J_theta Z(a0)' zeta,      zeta = wY[path] sqrt(dt) xi[path, step]   (adaptive process, no dL/dZ_sum)
// The checkpoint holds, per (tile, step), one row of 128 paths for every column of [a0 | h1 | h2] (RolloutParams::ckpt with
// ckpt_zeta == 0); zeta is regenerated from the Philox key.  A stage is 32 samples.  Per stage:
//     X [128 x 64] = [W2h_hi ; W2h_lo] [128 x s0] . [zeta_hi | zeta_lo]' [s0 x 64]   hidden MMA 1: tcgen05 SS, M = 128 = (hi | lo weights,
//                                                                    hidden slot), K = zeta column, N = (hi | lo, sample): all four
//                                                                    hi / lo products of 3xTF32 (+ lo.lo) in ONE pass of s0 / 8 MMAs
//     delta_2 = (X[r, 0..31] + X[r, 32..63] + X[64 + r, 0..31] + X[64 + r, 32..63]) * act'(h2)     r = h2 slot; epilogue warps:
//                                                                    TMEM -> registers (the W_lo rows through shared memory)
//                                                                    -> TMEM in place + a sample-major shared copy
//     X[h1 rows] += [W1h_hi ; W1h_lo] [128 x 32] . [delta_2_hi | delta_2_lo]'           hidden MMA 2 (the delta_2 rows of W1h are zero:
//                                                                    the MMA adds exact zeros to rows that already hold operands)
//     delta_1 = (the same four-term sum over the h1 slots) * act'(h1)
//     D0 [zeta col x act col]  += zeta'  . act      M = 128, N = s0 + 64, K = sample:  A = zeta' in TENSOR MEMORY, B = act rows (TMA, SW128)
//     D1 [delta row x act col] += delta' . act      M = 128 (64 used), N = s0 + 32, K = sample:  A = X itself: the epilogues leave
//                                                   hi(delta) in columns 0..31 and lo(delta) in columns 32..63 of rows 0..63
// Every product is FP32-equivalent (tc_sm100.cuh); hi = trunc_tf32(x) (what the tensor core keeps of a raw FP32 operand), lo = x - hi.
// Measured on this part (pspde_mma_probe): ONE tcgen05.mma costs its issuing thread >= 52 cycles whatever its shape
// (95 at N = 176), a commit -> mbarrier round trip ~ 800 cycles, and two issuing threads are slower than one.  Hence: as few MMA
// instructions as possible -- 41 per 32-sample stage: 13 + 4 hidden (hi | lo stacked along M and N), 24 weight-gradient -- and every
// buffer between two dependent MMAs double-buffered so that the round trips of neighbouring stages overlap.
//
// Where the operands live:
//   zeta    Philox -> shared, sample-major K-major tile zk[(col >> 2) * LBO + (sample | 32 lo) * 16 + (col & 3) * 4]: the B operand of
//           hidden MMA 1, one float4 per Philox call -> read back column-wise -> tensor memory (lane = zeta column, column = sample;
//           hi, lo): the A operand of dW0
//   act     TMA (CU_TENSOR_MAP_SWIZZLE_128B) -> shared [column][32 samples], a ring of three landing buffers (= the hi operand);
//           dead paths zeroed in place and the lo tile (two buffers) formed by 4 warps
//   delta   X (tensor memory, lane = row, two buffers) -> registers -> X in place and, delta_2 only, the sample-major shared tile dk
//           (two buffers)
//   weights [W2h_hi ; W2h_lo] (128 x s0), W2h = [W2[h2 rows]; W2[h1 rows]]; [W1h_hi ; W1h_lo] (128 x 32), W1h = [0; W1[h1 rows -> h2]]:
//           shared, K-major, once per CTA
//
// Warp roles (704 threads, one CTA per SM):
//   warps 0-7   epilogues of the hidden MMAs (lane quarter = warp & 3: q & 1 = delta_2 | delta_1 rows, q >= 2 = the W_lo partner rows;
//               16 samples each)
//   warps 8-15  zeta: Philox + Box-Muller (one warp = one 4-column group x 32 samples per call), then the tensor-memory copy
//   warps 16-19 activation rows: dead-path fix-up + lo tile; accumulator flush (quarter = warp & 3)
//   warp  20    TMA producer (one lane)
//   warp  21    MMA issuer (one lane), software-pipelined: hidden 2 (i) | hidden 1 (i + 1) | dW0(i + 1) | dW1(i)
#pragma once
#if !defined(PSPDE_EMULATE)
#include "grad_tc_kernels.cuh"

namespace pspde {

constexpr int kG2S = 32;                  // samples per stage
constexpr int kG2Sub = kCkP / kG2S;       // stages per (tile, step)
constexpr int kG2Threads = 22 * 32;
constexpr int kG2FlushStages = 16;        // accumulator flush period (the tensor core's FP32 accumulation is not round-to-nearest)
constexpr int kG2WGen = 8, kG2WLo = 16, kG2WTma = 20, kG2WMma = 21;
constexpr int kG2GenThreads = 256, kG2LoThreads = 128, kG2EpiThreads = 128, kG2EpiMain = 64;   // epilogue threads per hidden MMA (main + partner rows); those that finish a row
constexpr uint32_t kG2LboZ = 1040;        // bytes between 4-column groups of the sample-major tiles: (32 hi + 32 lo samples) x 16 B + 16
                                          // (LBO / 4 = 4 mod 32: the column-wise read-back and the scalar stores hit distinct banks)
constexpr uint32_t kG2LboW = 2048;        // weights: 128 rows ([hi (64) ; lo (64)] stacked along M) x 16 B per 4-column group

struct GradTc2Geom {
  int s0, act_rows, nA, nA1, dense, kz;
  uint32_t act_bytes;
  int nA1p;                                 // N of the dW1 MMA (M = 128: a multiple of 16 >= nA1)
  uint32_t o_act[3], o_lo[2], o_zk, o_dk[2], o_w2, o_w1, o_xb[2], o_bar, total;  // bytes from the 1 KB aligned base
  int c_d0, c_d1, c_a0[2], c_x[2];                                             // tensor-memory columns
};

inline bool grad_tc2_geom(const NetGeom& g, int s0, GradTc2Geom& t) {
  if (g.L != 3 || g.time_mode == TIME_NONE || g.seg_len[1] > 32 || g.seg_len[2] > 32 || (s0 & 7) || s0 < g.seg_len[0] || s0 > 128)
    return false;
  t.dense = g.kind == NET_DENSENET ? 1 : 0;
  t.s0 = s0; t.act_rows = s0 + 64; t.kz = s0 / 4;
  t.nA = (t.act_rows + 15) / 16 * 16;     // N of an M = 128 MMA is a multiple of 16
  t.nA1 = s0 + 32;                        // activation columns the hidden cotangents meet: [a0 | h1]
  t.nA1p = (t.nA1 + 15) / 16 * 16;
  if (t.nA > 256 || t.act_rows > 256) return false;
  int c = 0;
  t.c_d0 = c; c += t.nA;
  t.c_d1 = c; c += t.nA1p;
  t.c_a0[0] = c; c += kG2S; t.c_a0[1] = c; c += kG2S;
  t.c_x[0] = c; c += 2 * kG2S; t.c_x[1] = c; c += 2 * kG2S;
  if (c > 512) return false;
  // An MMA of N = nA reads nA - act_rows (< 16) rows past the data of a buffer: they alias the head of the next buffer (finite
  // values; those accumulator columns are never read), so the buffers are act_rows rows apart and only the last one is padded.
  t.act_bytes = (uint32_t)t.act_rows * 128u;    // a multiple of 1 KB (act_rows % 8 == 0)
  uint32_t o = 0;
  for (int s = 0; s < 3; ++s) { t.o_act[s] = o; o += t.act_bytes; }      // TMA landing ring = the hi operand
  for (int s = 0; s < 2; ++s) { t.o_lo[s] = o; o += t.act_bytes; }
  o += (uint32_t)(t.nA - t.act_rows) * 128u;
  t.o_zk = o; o += (uint32_t)t.kz * kG2LboZ;
  for (int h = 0; h < 2; ++h) { t.o_dk[h] = o; o += 8u * kG2LboZ; }
  for (int h = 0; h < 2; ++h) { t.o_xb[h] = o; o += 32u * 32u * 4u; }    // partner-row exchange of the two epilogue groups
  o = (o + 127u) & ~127u;
  t.o_w2 = o; o += (uint32_t)t.kz * kG2LboW;
  t.o_w1 = o; o += 8u * kG2LboW;
  t.o_bar = (o + 15u) & ~15u; o = t.o_bar + 32u * 8u;
  t.total = o + 1024u;
  return t.total <= 227u * 1024u;
}

// one RED.128 per lane: four consecutive floats (sm_90+); the SM retires ~1 RED warp instruction per 11 cycles whatever its width
__device__ __forceinline__ void g2_red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// index of (tile, activation column c, lane) in a CTA's partial: 4-column groups, [group][lane][4]
__host__ __device__ inline size_t g2_part_index(int nA, int tile, int c, int lane) {
  return ((size_t)(tile * (nA >> 2) + (c >> 2)) * 128 + (size_t)lane) * 4 + (size_t)(c & 3);
}

static __global__ void __launch_bounds__(kG2Threads, 1) grad_tc2_kernel(const __grid_constant__ CUtensorMap tmap, const RolloutParams prm,
                                                                        const GradTc2Geom tg, const int n_ts, const int flush_stages) {
  extern __shared__ float4 smem4_g2[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>(smem4_g2);
  __shared__ uint32_t tmem_base_s;
  const NetGeom& g = prm.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tg.o_bar);
  uint64_t* bar_full = bars + 16;       // [3] TMA bytes of the activation rows landed
  uint64_t* bar_free = bars + 19;       // [3] dW1 of the stage done: the landing buffer is free
  uint64_t* bar_lo = bars + 2;          // [2] dead paths zeroed, lo tile written
  uint64_t* bar_lofree = bars + 4;      // [2] dW1 of the stage done: the lo tile is free
  uint64_t* bar_zk = bars + 6;          // sample-major zeta tile written
  uint64_t* bar_a0 = bars + 7;          // zeta' in tensor memory written
  uint64_t* bar_d1 = bars + 8;          // hidden MMA 1 done (also: the sample-major zeta tile is free)
  uint64_t* bar_w0 = bars + 9;          // dW0 done: zeta' in tensor memory is free
  uint64_t* bar_e1 = bars + 10;         // delta_2 written (X in place + sample-major tile)
  uint64_t* bar_d2 = bars + 11;         // hidden MMA 2 done
  uint64_t* bar_e2 = bars + 12;         // delta_1 written
  uint64_t* bar_acc_full = bars + 13;   // D0 complete up to a flush point (committed right after dW0 of the flush stage)
  uint64_t* bar_acc_empty = bars + 14;  // D0 read out
  uint64_t* bar_acc1_full = bars + 22;  // D1 complete up to a flush point (after dW1)
  uint64_t* bar_acc1_empty = bars + 23; // D1 read out

  // ---- one-time setup
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int s = 0; s < 3; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_free[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&bar_lo[s], kG2LoThreads); tc::mbar_init(&bar_lofree[s], 1); }
    tc::mbar_init(bar_zk, kG2GenThreads); tc::mbar_init(bar_a0, kG2GenThreads);
    tc::mbar_init(bar_d1, 1); tc::mbar_init(bar_w0, 1); tc::mbar_init(bar_e1, kG2EpiMain); tc::mbar_init(bar_d2, 1);
    tc::mbar_init(bar_e2, kG2EpiMain);
    tc::mbar_init(bar_acc_full, 1); tc::mbar_init(bar_acc_empty, 4);
    tc::mbar_init(bar_acc1_full, 1); tc::mbar_init(bar_acc1_empty, 4);
    tc::mbar_fence_init();
  }
  if (tid == kG2WTma * 32) tc::tma_prefetch_desc(&tmap);
  for (uint32_t q = tid; q < tg.o_w2 / 16u; q += kG2Threads) reinterpret_cast<float4*>(smem)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  // weights of the hidden MMAs, K-major (hidden slot m, column k) at (k >> 2) * 1 KB + m * 16 + (k & 3) * 4, hi and lo.
  // slot m < 32: h2 column m (delta_2), else h1 column m - 32 (delta_1)
  {
    const LayerGeom& y2 = g.layer[2];
    const LayerGeom& y1 = g.layer[1];
    for (int q = tid; q < 64 * tg.s0; q += kG2Threads) {
      const int m = q & 63, k = q >> 6;
      const int sg = m < 32 ? 2 : 1, c = m & 31;
      float w = 0.f;
      if (c < g.seg_len[sg]) {
        const int lr = g.seg_off[sg] + c - y2.in_start;
        if (lr >= 0 && lr < y2.Kp) { const int idx = theta_index(g, 2, lr, k); if (idx >= 0) w = __ldg(prm.theta + idx); }
      }
      float hi, lo;
      tc::tf32_split_rn(w, hi, lo);
      const uint32_t off = (uint32_t)(k >> 2) * kG2LboW + (uint32_t)m * 16u + (uint32_t)(k & 3) * 4u;
      *reinterpret_cast<float*>(smem + tg.o_w2 + off) = hi;                  // rows 0..63: hi, rows 64..127: lo
      *reinterpret_cast<float*>(smem + tg.o_w2 + off + 64u * 16u) = lo;
    }
    for (int q = tid; q < 64 * 32; q += kG2Threads) {
      const int m = q & 63, k = q >> 6;           // k = h2 column (delta_2 slot)
      float w = 0.f;
      if (m >= 32 && (m & 31) < g.seg_len[1]) {
        const int lr = g.seg_off[1] + (m & 31) - y1.in_start;
        if (lr >= 0 && lr < y1.Kp) { const int idx = theta_index(g, 1, lr, k); if (idx >= 0) w = __ldg(prm.theta + idx); }
      }
      float hi, lo;
      tc::tf32_split_rn(w, hi, lo);
      const uint32_t off = (uint32_t)(k >> 2) * kG2LboW + (uint32_t)m * 16u + (uint32_t)(k & 3) * 4u;
      *reinterpret_cast<float*>(smem + tg.o_w1 + off) = hi;
      *reinterpret_cast<float*>(smem + tg.o_w1 + off + 64u * 16u) = lo;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  const int my_ts = (n_ts - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // (tile, step) pairs of this CTA
  const int n_it = my_ts > 0 ? my_ts * kG2Sub : 0;
  const bool unit = prm.ckpt_unit != 0;
  // The CTAs flush their accumulators at DIFFERENT stages (offset by the CTA index): 148 CTAs flushing in the same few
  // microseconds saturate the L2's atomic units (measured: ~14 k cycles per flush when synchronous).
  const int flush_off = (int)(blockIdx.x % (unsigned)flush_stages);
  // first local path index and step of stage iteration `it`
  auto stage_path0 = [&](int it) {
    const int ts = (int)blockIdx.x + (it / kG2Sub) * (int)gridDim.x;
    return (prm.tile0 + ts / prm.N) * kCkP + (it % kG2Sub) * kG2S;
  };
  auto stage_step = [&](int it) { return (unsigned)(((int)blockIdx.x + (it / kG2Sub) * (int)gridDim.x) % prm.N); };

  if (warp < kG2WGen) {
    // =============================================================== epilogues of the hidden MMAs
    // X rows (= tensor-memory lanes, M = 128): 0..31 delta_2 slots, 32..63 delta_1 slots -- products with W_hi -- and
    // 64..127 the same slots' products with W_lo (the "partner" rows).  Quarter q of the lanes: q & 1 = which delta,
    // q >= 2 = partner.  A partner warp only adds its hi | lo partial sums and hands them over through shared memory.
    const int q = warp & 3, sh = warp >> 2;          // lane quarter; 16-sample half
    const bool is_d2 = (q & 1) == 0, partner = q >= 2;
    const int c = lane;                              // column inside the hidden segment
    const int sg = is_d2 ? 2 : 1;
    const bool live = c < g.dims[sg];
    const int h_row = tg.s0 + (is_d2 ? 32 : 0) + c;  // tile row of the hidden activation
    const uint32_t lane_addr = ((uint32_t)(32 * q)) << 16;
    float* xb = reinterpret_cast<float*>(smem + tg.o_xb[is_d2 ? 0 : 1]) + (16 * sh) * 32 + lane;     // [sample][slot]
    // debug (CTA 0): warp 0 [0] wait hidden MMA 1, [1] wait lo, [3] delta_2; warp 1 [4] waits, [5] delta_1
    PhaseTimer pt_;
    pt_.start(prm.prof, (lane == 0 && (warp == 0 || warp == 1)) ? 0 : 1);
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const uint32_t p1 = (uint32_t)it & 1u, p2 = (uint32_t)(it >> 1) & 1u;
      tc::mbar_wait(is_d2 ? bar_d1 : bar_d2, p1);
      pt_.mark(warp == 0 ? 0 : 4);
      if (!partner) tc::mbar_wait(&bar_lo[s], p2);   // the hidden activations of dead paths are zero from here on
      pt_.mark(warp == 0 ? 1 : 4);
      tc::fence_after_sync();
      const uint32_t xa = tbase + lane_addr + (uint32_t)tg.c_x[s] + 16u * (uint32_t)sh;
      float P[16], Q[16];
      tc::tmem_ld16(xa, P);                          // . x_hi partial sums
      tc::tmem_ld16(xa + 32u, Q);                    // . x_lo partial sums
      if (partner) {
        tc::wait_ld();
#pragma unroll
        for (int n = 0; n < 16; ++n) xb[32 * n] = P[n] + Q[n];
        tc::fence_before_sync();
        gt_named_bar(is_d2 ? 6 : 7, kG2EpiThreads);
        continue;
      }
      const uint8_t* tH = smem + tg.o_act[it % 3];
      float4 h4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h4[j] = *reinterpret_cast<const float4*>(tH + gt_swz(h_row, 4 * sh + j));
      tc::wait_ld();
      gt_named_bar(is_d2 ? 6 : 7, kG2EpiThreads);    // the partner rows' sums are in shared memory
      if (warp == 0) pt_.mark(24);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hv[4] = {h4[j].x, h4[j].y, h4[j].z, h4[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // act': relu(.)^2 -> 2 relu(pre) = 2 sqrt(h) (sqrt.approx: 1 ulp); tanh -> 1 - h^2
          const float pre = (P[4 * j + i] + Q[4 * j + i]) + xb[32 * (4 * j + i)];
          const float dv = live ? pre * (tg.dense ? 2.0f * gt_sqrt_approx(hv[i]) : (1.0f - hv[i] * hv[i])) : 0.f;
          P[4 * j + i] = trunc_tf32(dv);                        // P <- hi (what the tensor core would keep of dv anyway),
          Q[4 * j + i] = dv - P[4 * j + i];                     // Q <- lo (exact; its own truncation loses < 2^-21 |dv|)
        }
      }
      if (warp == 0) pt_.mark(25);
      tc::tmem_st16(xa, P);                          // in place: the A operand of dW1
      tc::tmem_st16(xa + 32u, Q);
      if (is_d2) {                                   // sample-major copy [hi | lo]: B operand of hidden MMA 2
        uint8_t* dk = smem + tg.o_dk[s] + (uint32_t)(c >> 2) * kG2LboZ + (uint32_t)(c & 3) * 4u + (uint32_t)(16 * sh) * 16u;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          *reinterpret_cast<float*>(dk + 16 * n) = P[n];
          *reinterpret_cast<float*>(dk + 512 + 16 * n) = Q[n];
        }
        if (warp == 0) pt_.mark(26);
        tc::fence_proxy_async();
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(is_d2 ? bar_e1 : bar_e2);
      pt_.mark(warp == 0 ? 3 : 5);
    }
  } else if (warp < kG2WLo) {
    // =============================================================== zeta: Philox -> shared tile -> tensor memory
    const int gw = warp - kG2WGen;                   // 0..7
    const int q = warp & 3, half = gw >> 2;          // tensor-memory copy: lane quarter (zeta columns 32 q + lane), 16-sample half
    const uint32_t lane_addr = ((uint32_t)(32 * q)) << 16;
    const float sqdt = sqrtf(prm.dt);
    const int zc = 32 * q + lane;                    // zeta column of this lane in the tensor-memory copy
    const uint8_t* zrd = smem + tg.o_zk + (uint32_t)(zc >> 2) * kG2LboZ + (uint32_t)(zc & 3) * 4u + (uint32_t)(16 * half) * 16u;
    const bool zin = zc < tg.s0;
    // debug (CTA 0, warp 8): [6] Philox, [7] wait sample-major tile free, [8] its stores + barrier, [9] read-back, [10] wait A0 free + st
    PhaseTimer pt_;
    pt_.start(prm.prof, (lane == 0 && warp == kG2WGen) ? 0 : 1);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t p1 = (uint32_t)it & 1u;
      const int kl = stage_path0(it) + lane;         // Philox: lane = sample, warp gw draws groups gw, gw + 8, ...
      const unsigned nstep = stage_step(it);
      const float wk = (kl < prm.K_local) ? __ldg(prm.wY + kl) : 0.f;
      const float sc = wk * sqdt;
      float4 zh[4], zl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gq = gw + 8 * i;
        float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * gq < prm.d) e4 = philox_normal4((unsigned)(prm.k_offset + kl), nstep, (unsigned)gq, prm.offset, prm.seed);
        // wk == 0 (padding, dropped trajectory): exact zeros whatever the draw
        const float z0 = wk != 0.f ? sc * e4.x : 0.f;
        const float z1 = (wk != 0.f && 4 * gq + 1 < prm.d) ? sc * e4.y : 0.f;
        const float z2 = (wk != 0.f && 4 * gq + 2 < prm.d) ? sc * e4.z : 0.f;
        const float z3 = (wk != 0.f && 4 * gq + 3 < prm.d) ? sc * e4.w : 0.f;
        // hi = what the tensor core keeps of z (truncation), lo = z - hi exactly (its own truncation loses < 2^-21 |z|)
        zh[i].x = trunc_tf32(z0); zl[i].x = z0 - zh[i].x; zh[i].y = trunc_tf32(z1); zl[i].y = z1 - zh[i].y;
        zh[i].z = trunc_tf32(z2); zl[i].z = z2 - zh[i].z; zh[i].w = trunc_tf32(z3); zl[i].w = z3 - zh[i].w;
      }
      pt_.mark(6);
      tc::mbar_wait(bar_d1, p1 ^ 1u);                // hidden MMA 1 of the previous stage has read the tile
      pt_.mark(7);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gq = gw + 8 * i;
        if (gq < tg.kz) {
          uint8_t* dst = smem + tg.o_zk + (uint32_t)gq * kG2LboZ + (uint32_t)lane * 16u;
          *reinterpret_cast<float4*>(dst) = zh[i];
          *reinterpret_cast<float4*>(dst + 512) = zl[i];
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(bar_zk);
      gt_named_bar(4, kG2GenThreads);                // all 26 groups are in shared memory
      pt_.mark(8);
      float hi[16], lo[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        hi[n] = zin ? *reinterpret_cast<const float*>(zrd + 16 * n) : 0.f;
        lo[n] = zin ? *reinterpret_cast<const float*>(zrd + 512 + 16 * n) : 0.f;
      }
      gt_named_bar(5, kG2GenThreads);                // every warp has its columns: the tile may be overwritten by the next stage's draws
      pt_.mark(9);
      tc::mbar_wait(bar_w0, p1 ^ 1u);                // dW0 of the previous stage read zeta' from tensor memory
      tc::tmem_st16(tbase + lane_addr + (uint32_t)tg.c_a0[0] + 16u * half, hi);
      tc::tmem_st16(tbase + lane_addr + (uint32_t)tg.c_a0[1] + 16u * half, lo);
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(bar_a0);
      pt_.mark(10);
    }
  } else if (warp < kG2WTma) {
    // =============================================================== activation rows: fix-up + lo tile; accumulator flush
    const int t = tid - kG2WLo * 32;                 // 0..127
    const int qtr = warp & 3;
    const int pos = t & 7, r0 = t >> 3;              // chunk position; rows r0 + 16 i keep (row & 7), hence the sample quad
    const int j = pos ^ (r0 & 7);
    uint32_t n_flush = 0, n_flush1 = 0;
    bool d1_pending = false;
    float* gp = prm.grad_partial + (size_t)blockIdx.x * (2 * 128 * tg.nA);
    const uint32_t la = ((uint32_t)(32 * qtr)) << 16;
    // 32 columns per tensor-memory round trip (the load latency, not the REDs, is what a flush costs); pad columns / lanes skipped
    auto flush_cols = [&](int tile, int ccol, int c_begin, int c_end, bool on) {
      int c0 = c_begin;
      for (; c0 + 32 <= c_end; c0 += 32) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) tc::tmem_ld8(tbase + la + (uint32_t)(ccol + c0 + 8 * u), v[u]);
        tc::wait_ld();
        if (on) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            g2_red_add4(gp + g2_part_index(tg.nA, tile, c0 + 8 * u, 32 * qtr + lane), v[u][0], v[u][1], v[u][2], v[u][3]);
            g2_red_add4(gp + g2_part_index(tg.nA, tile, c0 + 8 * u + 4, 32 * qtr + lane), v[u][4], v[u][5], v[u][6], v[u][7]);
          }
        }
      }
      for (; c0 < c_end; c0 += 8) {
        float v[8];
        tc::tmem_ld8(tbase + la + (uint32_t)(ccol + c0), v);
        tc::wait_ld();
        if (on) {
          g2_red_add4(gp + g2_part_index(tg.nA, tile, c0, 32 * qtr + lane), v[0], v[1], v[2], v[3]);
          g2_red_add4(gp + g2_part_index(tg.nA, tile, c0 + 4, 32 * qtr + lane), v[4], v[5], v[6], v[7]);
        }
      }
    };
    auto flush_d1 = [&]() {
      tc::mbar_wait(bar_acc1_full, n_flush1 & 1u);
      tc::fence_after_sync();
      flush_cols(1, tg.c_d1, 0, tg.nA1, qtr < 2);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_acc1_empty);
      ++n_flush1;
    };
    PhaseTimer pt_;      // debug (CTA 0, warp 16): [11] wait TMA, [12] fix-up + lo pass, [13] flush
    pt_.start(prm.prof, (lane == 0 && warp == kG2WLo) ? 0 : 1);
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const uint32_t p2 = (uint32_t)(it >> 1) & 1u;
      const int s3 = it % 3;
      const uint32_t p3 = (uint32_t)(it / 3) & 1u;
      uint8_t* tH = smem + tg.o_act[s3];
      uint8_t* tL = smem + tg.o_lo[s];
      bool keep[4] = {true, true, true, true};
      if (unit) {                                    // forward-written rows: drop every row of a path whose cotangent is zero
        const int kq = stage_path0(it) + 4 * j;
#pragma unroll
        for (int i = 0; i < 4; ++i) keep[i] = (kq + i < prm.K_local) && __ldg(prm.wY + kq + i) != 0.f;
      }
      tc::mbar_wait(&bar_full[s3], p3);
      tc::mbar_wait(&bar_lofree[s], p2 ^ 1u);        // dW1 of stage it - 2 has read this lo tile (first use passes)
      pt_.mark(11);
      // rows r0 + 16 i: four loads in flight at a time (the shared-memory latency under the tensor core's operand traffic is
      // several hundred cycles)
      for (int rb = r0; rb < tg.act_rows; rb += 64) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rb + 16 * i;
          if (r < tg.act_rows) v[i] = *reinterpret_cast<const float4*>(tH + (uint32_t)r * 128u + (uint32_t)(pos << 4));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rb + 16 * i;
          if (r < tg.act_rows) {
            const uint32_t off = (uint32_t)r * 128u + (uint32_t)(pos << 4);
            float4 w = v[i];
            if (unit) {
              w.x = keep[0] ? w.x : 0.f; w.y = keep[1] ? w.y : 0.f; w.z = keep[2] ? w.z : 0.f; w.w = keep[3] ? w.w : 0.f;
              *reinterpret_cast<float4*>(tH + off) = w;
            }
            *reinterpret_cast<float4*>(tL + off) = lo4(w);
          }
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&bar_lo[s]);
      pt_.mark(12);
      // Accumulator flush: raw accumulators -> this CTA's partial (RED.128, one writer per address, L2 resident).  D0 is complete
      // as soon as dW0 of the flush stage is; D1 only after dW1, which the issuer places AFTER dW0 of the next stage -- and that
      // one needs the next lo tile from these warps.  So: D0 right after the lo pass of the flush stage, D1 after the lo pass
      // of the stage that follows (or after the loop).
      if (d1_pending) { flush_d1(); d1_pending = false; pt_.mark(13); }
      const bool flush_now = ((it + 1 + flush_off) % flush_stages == 0) || it == n_it - 1;
      if (flush_now) {
        tc::mbar_wait(bar_acc_full, n_flush & 1u);
        tc::fence_after_sync();
        flush_cols(0, tg.c_d0, 0, tg.act_rows, 32 * qtr + lane < tg.s0);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_acc_empty);
        ++n_flush;
        d1_pending = true;
        pt_.mark(13);
      }
    }
    if (d1_pending) flush_d1();
  } else if (warp == kG2WTma) {
    // =============================================================== TMA producer
    if (lane == 0) {
      PhaseTimer pt_;    // debug (CTA 0): [14] wait for a free buffer
      pt_.start(prm.prof, 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % 3;
        pt_.mark(15);
        tc::mbar_wait(&bar_free[s], ((uint32_t)(it / 3) & 1u) ^ 1u);       // first use of a buffer passes immediately
        pt_.mark(14);
        const int ts = (int)blockIdx.x + (it / kG2Sub) * (int)gridDim.x, sub = it % kG2Sub;
        tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)tg.act_rows * 128u);
        tc::tma_load_3d(smem + tg.o_act[s], &tmap, &bar_full[s], sub * kG2S, 0, ts);
      }
    }
  } else {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t sb = tc::smem_u32(smem);
      const uint32_t id_h = tc::idesc_tf32(128, 2 * kG2S);
      const uint32_t id_w0 = tc::idesc_tf32(128, tg.nA), id_w1 = tc::idesc_tf32(128, tg.nA1p);
      // descriptors: only the 14-bit start address (>> 4) changes from one MMA to the next -> add to the low word
      const uint64_t dw = tc::smem_desc(sb + tg.o_w2, kG2LboW, 128u), dv = tc::smem_desc(sb + tg.o_w1, kG2LboW, 128u);
      const uint64_t dz = tc::smem_desc(sb + tg.o_zk, kG2LboZ, 128u);
      const uint64_t dd0 = tc::smem_desc(sb + tg.o_dk[0], kG2LboZ, 128u), dd1 = tc::smem_desc(sb + tg.o_dk[1], kG2LboZ, 128u);
      const uint64_t dah0 = tc::smem_desc_sw128(sb + tg.o_act[0]), dah1 = tc::smem_desc_sw128(sb + tg.o_act[1]);
      const uint64_t dah2 = tc::smem_desc_sw128(sb + tg.o_act[2]);
      const uint64_t dal0 = tc::smem_desc_sw128(sb + tg.o_lo[0]), dal1 = tc::smem_desc_sw128(sb + tg.o_lo[1]);
      const uint32_t cx0 = (uint32_t)tg.c_x[0], cx1 = (uint32_t)tg.c_x[1];
      constexpr uint64_t kStepW = (2u * kG2LboW) >> 4, kStepZ = (2u * kG2LboZ) >> 4, kStepA = 32u >> 4;
      // X (+)= [W_hi ; W_lo] . [x_hi | x_lo]'  (M = 128, N = 64): all four hi / lo products in ONE pass; the epilogue adds
      // X[r, 0..31] + X[r, 32..63] + X[64 + r, 0..31] + X[64 + r, 32..63] (lo . lo, 2^-22 relative, comes along for free)
      auto hidden_mma = [&](uint64_t w, uint64_t x, int nk8, uint32_t dcol, bool acc0) {
        if (nk8 == 13) {                             // the C2 / C5 shape (s0 = 104): fully unrolled, descriptor offsets are immediates
#pragma unroll
          for (int ks = 0; ks < 13; ++ks)
            tc::mma_tf32_ss(tbase + dcol, w + (uint64_t)ks * kStepW, x + (uint64_t)ks * kStepZ, id_h, acc0 || ks > 0);
        } else if (nk8 == 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc::mma_tf32_ss(tbase + dcol, w + (uint64_t)ks * kStepW, x + (uint64_t)ks * kStepZ, id_h, acc0 || ks > 0);
        } else {
          for (int ks = 0; ks < nk8; ++ks)
            tc::mma_tf32_ss(tbase + dcol, w + (uint64_t)ks * kStepW, x + (uint64_t)ks * kStepZ, id_h, acc0 || ks > 0);
        }
      };
      // D (+)= A' (tensor memory: hi at column ca, lo at ca + 32) . act rows of buffer s over the 32 samples of the stage
      auto wgrad_mma = [&](uint32_t ca, int s, int s3, uint32_t dcol, uint32_t idesc, bool acc0) {
        const uint64_t bh = s3 == 0 ? dah0 : s3 == 1 ? dah1 : dah2, bl = s ? dal1 : dal0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {       // small terms first
          const uint32_t a = tbase + ca + (pass == 0 ? 32u : 0u);
          const uint64_t b = pass == 1 ? bl : bh;
#pragma unroll
          for (int ks = 0; ks < kG2S / 8; ++ks)
            tc::mma_tf32_ts(tbase + dcol, a + 8u * (uint32_t)ks, b + (uint64_t)ks * kStepA, idesc, acc0 || pass > 0 || ks > 0);
        }
      };
      uint32_t nf0 = 0, nf1 = 0;                      // flushes of D0 / D1 committed so far
      // debug (CTA 0): [16] wait zeta' + lo (+ flush), [17] dW0, [18] wait delta_2, [19] hidden MMA 2, [20] wait zeta tile,
      // [21] hidden MMA 1, [22] wait delta_1, [23] dW1 + commits
      PhaseTimer pt_;
      pt_.start(prm.prof, 0);
      auto is_first = [&](int j) { return j == 0 || ((j + flush_off) % flush_stages) == 0; };   // accumulators start over after a flush
      auto is_flush = [&](int j) { return ((j + 1 + flush_off) % flush_stages == 0) || j == n_it - 1; };
      // dW0(j): zeta' . act
      auto issue_dw0 = [&](int j) {
        const bool first = is_first(j);
        if (first && nf0 > 0) tc::mbar_wait(bar_acc_empty, (nf0 - 1u) & 1u);          // D0 has been read out
        tc::mbar_wait(bar_a0, (uint32_t)j & 1u);
        tc::mbar_wait(&bar_lo[j & 1], (uint32_t)(j >> 1) & 1u);
        pt_.mark(16);
        tc::fence_after_sync();
        wgrad_mma((uint32_t)tg.c_a0[0], j & 1, j % 3, (uint32_t)tg.c_d0, id_w0, !first);
        tc::mma_commit(bar_w0);
        if (is_flush(j)) { tc::mma_commit(bar_acc_full); ++nf0; }
        pt_.mark(17);
      };
      // dW1(j): [delta_2 | delta_1]' . act
      auto issue_dw1 = [&](int j) {
        const bool first = is_first(j);
        tc::mbar_wait(bar_e2, (uint32_t)j & 1u);
        if (first && nf1 > 0) tc::mbar_wait(bar_acc1_empty, (nf1 - 1u) & 1u);         // D1 has been read out
        pt_.mark(22);
        tc::fence_after_sync();
        wgrad_mma((j & 1) ? cx1 : cx0, j & 1, j % 3, (uint32_t)tg.c_d1, id_w1, !first);
        tc::mma_commit(&bar_free[j % 3]);
        tc::mma_commit(&bar_lofree[j & 1]);
        if (is_flush(j)) { tc::mma_commit(bar_acc1_full); ++nf1; }
        pt_.mark(23);
      };
      if (n_it > 0) {                                 // prologue: hidden MMA 1 and dW0 of stage 0
        tc::mbar_wait(bar_zk, 0u);
        tc::fence_after_sync();
        hidden_mma(dw, dz, tg.kz / 2, cx0, false);
        tc::mma_commit(bar_d1);
        issue_dw0(0);
      }
      // steady state: hidden 2 (it) | hidden 1 (it + 1) | dW0(it + 1) | dW1(it) -- the delta_1 epilogue of stage it (the longest
      // dependent hop) runs under hidden 1 and dW0 of the next stage
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t p1 = (uint32_t)it & 1u;
        tc::mbar_wait(bar_e1, p1);
        pt_.mark(18);
        tc::fence_after_sync();
        hidden_mma(dv, s ? dd1 : dd0, 4, s ? cx1 : cx0, true);
        tc::mma_commit(bar_d2);
        pt_.mark(19);
        if (it + 1 < n_it) {
          tc::mbar_wait(bar_zk, p1 ^ 1u);
          pt_.mark(20);
          tc::fence_after_sync();
          hidden_mma(dw, dz, tg.kz / 2, s ? cx0 : cx1, false);
          tc::mma_commit(bar_d1);
          pt_.mark(21);
          issue_dw0(it + 1);
        }
        issue_dw1(it);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// partial[cta][g2_part_index(tile, activation column m, lane)] -> grad_theta.
//   tile 0: lane <-> zeta column n: W2[activation column m][n]
//   tile 1: lane r < 64 <-> hidden slot r: r < 32 -> W1[m][r] (delta_2), else W0[m][r - 32] (delta_1)
// activation column m: checkpoint columns [a0 (s0) | h1 (32) | h2 (32)].  Fixed summation order, fp64.
static __global__ void reduce_grad_tc2_kernel(const NetGeom g, const GradTc2Geom tg, const float* __restrict__ partial, int nparts,
                                              float* __restrict__ out) {
  const int per = 2 * 128 * tg.nA;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
    // q = g2_part_index(nA, tile, m, lane)
    const int lane = (q >> 2) & 127, cg = q >> 9, tile = cg / (tg.nA >> 2), m = 4 * (cg - tile * (tg.nA >> 2)) + (q & 3);
    int col = -1;
    if (m < tg.s0) { if (m < g.seg_len[0]) col = m; }
    else if (m < tg.s0 + 32) { if (m - tg.s0 < g.seg_len[1]) col = g.seg_off[1] + (m - tg.s0); }
    else if (m < tg.s0 + 64) { if (m - tg.s0 - 32 < g.seg_len[2]) col = g.seg_off[2] + (m - tg.s0 - 32); }
    if (col < 0) continue;
    int l, n;
    if (tile == 0) { l = 2; n = lane; }
    else {
      if (lane >= 64 || m >= tg.nA1) continue;
      const int r = lane;
      if (r < 32) { l = 1; n = r; } else { l = 0; n = r - 32; }
    }
    const LayerGeom& y = g.layer[l];
    const int r = col - y.in_start;
    if (r < 0 || r >= y.Kp || n >= y.N) continue;
    const int idx = theta_index(g, l, r, n);
    if (idx < 0) continue;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partial[(size_t)p * per + q];
    out[idx] = (float)s;
  }
}

}  // namespace pspde
#endif  // !PSPDE_EMULATE
