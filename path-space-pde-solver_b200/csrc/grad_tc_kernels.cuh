// grad_tc_kernels.cuh -- gradient accumulation from the checkpoint rows: a TMA-fed, warp-specialised tcgen05 pipeline.
// Same contract as grad_kernel (grad_kernels.cuh): a pure streaming kernel over (trajectory, step) samples,
//     dtheta = sum_samples J_theta Z(a0)' zeta                                   (reference: loss.backward(), solver.py:221).
//
// The reduction dimension K of the weight gradient
//     D[act col][cot col] += sum_samples act[sample][act col] * cot[sample][cot col]
// is the SAMPLE index, M = activation column, N = cotangent column.  The checkpoint keeps, per (tile, step), one row of
// 128 consecutive paths for every column of [a0 | h1 | h2 | zeta] (RolloutParams::ckpt: column-major, the path index
// contiguous), i.e. exactly the K-major operand form of tcgen05.mma with K = sample.  A stage is 32 samples -- one 128-byte
// row per column -- and arrives by TMA (cp.async.bulk.tensor, CU_TENSOR_MAP_SWIZZLE_128B) straight in the operand layout
// (8 rows x 128 B atoms, 16-byte chunk index xor-ed with the row inside the atom): no transposing copy, no register staging.
// All columns -- [a0 | h1 | h2] and [zeta | delta_2 | delta_1] -- are rows of ONE such tile, so the A operand (M tiles: rows
// 0..127, 128..255) and the B operand (rows from zeta on, N = 176 at the C2 shape) are windows of the same tile.
// FP32 equivalence as in tc_sm100.cuh: the tile holds raw FP32 values, which the tensor core truncates to TF32
// (hi = trunc(x)); a second tile holds lo = rna_tf32(x - trunc(x)); three passes lo*hi + hi*lo + hi*hi into the FP32
// accumulators in tensor memory (n_mt x nB columns), flushed to the CTA's partial every few stages (see below).
//
// Warp roles (640 threads, one CTA per SM, 2-stage ring of (hi, lo) tiles = 172 KB at the C2 shape):
//   warps 0-15  hidden cotangents: delta_2 = (zeta W2[h2 rows]') act'(h2), delta_1 = (zeta W2[h1 rows]' + delta_2 W1[h1 rows]')
//               act'(h1) in FP32 FMA (warp = 4 hidden columns; lane = 4 samples x a quarter of the reduction range, operands
//               read from the swizzled tile), appended to the tile as rows (hi and lo)
//   warp  16    TMA producer (one lane): waits for a free stage, arms the mbarrier, issues the two box loads
//   warp  17    MMA issuer (one lane): when the lo tile and the delta rows of a stage exist; tcgen05.commit frees the stage
//   warps 18-19 lo tile of the loaded rows (element-wise, layout agnostic)
//   warps 16-19 additionally flush the accumulators (one lane quarter of tensor memory each)
#pragma once
#if !defined(PSPDE_EMULATE)
#include <cuda.h>
#include "grad_kernels.cuh"
#include "tc_sm100.cuh"

namespace pspde {

constexpr int kGtS = 32;                 // samples per stage = one 128-byte swizzle row of K = 4 MMA k steps
constexpr int kGtStages = 2;
constexpr int kGtWorkers = 512;          // hidden-cotangent threads (warps 0-15)
constexpr int kGtThreads = 640;
constexpr int kGtUtil = kGtWorkers / 32; // first utility warp: TMA producer; +1 MMA issuer; +2, +3 lo tile
constexpr int kGtSub = kCkP / kGtS;      // stages per (tile, step)
constexpr int kGtFlushStages = 16;       // accumulator flush period: 16 stages x 12 MMAs per accumulator

struct GradTcGeom {
  int s0;              // columns of a0 in the checkpoint (multiple of 8)
  int cols;            // checkpoint columns per sample: [a0 (s0) | h1 (32) | h2 (32) | zeta (s0)]
  int act_rows;        // s0 + 64: rows [a0 | h1 | h2] of the operand tile
  int r_ze, r_d2, r_d1, rows;   // first row of zeta / delta_2 / delta_1; tile rows (multiple of 8, >= 128 n_mt)
  int box_rows;        // rows per TMA box (cols / 2)
  int nB;              // N of the weight-gradient product: zeta (padded to nE) | delta_2 (32) | delta_1 (32)
  int nE;              // zeta columns padded to a multiple of 16: their MMAs are issued before the hidden cotangents exist
  int dense;
  int n_mt;            // M tiles of 128 activation columns (1 when [a0 | h1 | h2] has at most 128 columns)
  // compact weights for the hidden cotangents, k4-blocked: W2c rows = [h1 (32) | h2 (32)] columns (zero rows where
  // the layer does not read the column), W1c rows = h1 (32) columns
  int w2_nng, w1_nng, o_w2, o_w1;
  uint32_t tile_bytes;                   // rows * 128
  uint32_t o_hi[kGtStages], o_lo[kGtStages], o_w, o_bar, total;   // shared-memory byte offsets from the 1 KB aligned base
};

inline bool grad_tc_geom(const NetGeom& g, int d, int s0, GradTcGeom& t) {
  if (g.L != 3 || g.time_mode == TIME_NONE || g.seg_len[1] > 32 || g.seg_len[2] > 32 || (s0 & 7) || s0 < g.seg_len[0]) return false;
  (void)d;
  t.dense = g.kind == NET_DENSENET ? 1 : 0;
  t.s0 = s0;
  t.cols = 2 * s0 + 64;
  t.act_rows = s0 + 64;
  t.box_rows = t.cols / 2;               // s0 + 32: a multiple of 8, so the second box starts on an atom boundary
  if (t.box_rows > 256) return false;
  t.nE = ((s0 + 15) / 16) * 16;
  t.nB = t.nE + 64;
  t.r_ze = t.act_rows; t.r_d2 = t.r_ze + t.nE; t.r_d1 = t.r_d2 + 32;
  t.n_mt = t.act_rows > 128 ? 2 : 1;
  t.rows = t.r_d1 + 32;
  if (t.rows < 128 * t.n_mt) t.rows = 128 * t.n_mt;          // M tile mt reads rows [128 mt, 128 mt + 128)
  if (t.nB > 256 || t.n_mt * t.nB > 512) return false;
  t.w2_nng = g.layer[2].nng; t.w1_nng = g.layer[1].nng;
  t.tile_bytes = (uint32_t)t.rows * 128u;                    // a multiple of 1 KB (rows % 8 == 0)
  uint32_t o = 0;
  for (int s = 0; s < kGtStages; ++s) { t.o_hi[s] = o; o += t.tile_bytes; t.o_lo[s] = o; o += t.tile_bytes; }
  t.o_w = o;
  t.o_w2 = 0; t.o_w1 = 64 * t.w2_nng * 4;
  o += (uint32_t)(64 * t.w2_nng * 4 + 32 * t.w1_nng * 4) * 4u;
  t.o_bar = (o + 15u) & ~15u; o = t.o_bar + 16u * 8u;
  t.total = o + 1024u;                                       // slack for the 1 KB alignment of the base
  return t.total <= 227u * 1024u;
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// lo = x - trunc(x), rounded to TF32 so that the tensor core's own truncation of it is exact
__device__ __forceinline__ float lo1(float x) { return tc::tf32_hi(x - trunc_tf32(x)); }
__device__ __forceinline__ float4 lo4(const float4& v) { return make_float4(lo1(v.x), lo1(v.y), lo1(v.z), lo1(v.w)); }

// byte offset of (row r, sample quad j) in a swizzled tile: 16-byte chunk index xor-ed with the row inside the 8-row atom
__device__ __forceinline__ uint32_t gt_swz(int r, int j) { return (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4); }

__device__ __forceinline__ float gt_sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void gt_named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// n_ts = number of (tile, step) pairs in the checkpoint; CTA b streams pairs b, b + grid, ... (4 stages each).
// flush_stages: accumulator flush period in stages.
static __global__ void __launch_bounds__(kGtThreads, 1) grad_tc_kernel(const __grid_constant__ CUtensorMap tmap, const RolloutParams prm,
                                                                const GradTcGeom tg, const int n_ts, const int flush_stages) {
  extern __shared__ float4 smem4_gt[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>(smem4_gt);
  __shared__ uint32_t tmem_base_s;
  const NetGeom& g = prm.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sW = reinterpret_cast<float*>(smem + tg.o_w);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tg.o_bar);
  uint64_t* bar_full = bars;                  // [stage] TMA bytes landed
  uint64_t* bar_empty = bars + 2;             // [stage] the tensor core is done with the stage
  uint64_t* bar_lo = bars + 4;                // [stage] lo tile of the loaded rows written
  uint64_t* bar_dl = bars + 6;                // [stage] delta rows (hi, lo) written
  uint64_t* bar_acc_full = bars + 8;          // accumulators complete up to a flush point
  uint64_t* bar_acc_empty = bars + 9;         // accumulators read out: the next MMA may overwrite them

  // ---- one-time setup
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    for (int s = 0; s < kGtStages; ++s) {
      tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1);
      tc::mbar_init(&bar_lo[s], 64); tc::mbar_init(&bar_dl[s], kGtWorkers);
    }
    tc::mbar_init(bar_acc_full, 1); tc::mbar_init(bar_acc_empty, 4);
    tc::mbar_fence_init();
  }
  if (tid == kGtUtil * 32) tc::tma_prefetch_desc(&tmap);
  for (uint32_t q = tid; q < (uint32_t)kGtStages * 2u * tg.tile_bytes / 16u; q += kGtThreads)
    reinterpret_cast<float4*>(smem)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 1; l <= 2; ++l) {            // compact k4-blocked weights: row = hidden column slot (32 per segment)
    const LayerGeom& y = g.layer[l];
    const int rows = l == 2 ? 64 : 32;
    float* dst = sW + (l == 2 ? tg.o_w2 : tg.o_w1);
    for (int q = tid; q < rows * y.Np; q += kGtThreads) {
      const int blk = q >> 4, in = q & 15;
      const int r = 4 * (blk / y.nng) + (in >> 2), n = 4 * (blk % y.nng) + (in & 3);
      const int sg = 1 + (r >> 5), c = r & 31;                 // segment and column inside it
      int idx = -1;
      if (c < g.seg_len[sg]) {
        const int lr = g.seg_off[sg] + c - y.in_start;         // row of W_l (relative to the first column it reads)
        if (lr >= 0 && lr < y.Kp) idx = theta_index(g, l, lr, n);
      }
      dst[q] = idx >= 0 ? __ldg(prm.theta + idx) : 0.f;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;

  const int my_ts = (n_ts - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // (tile, step) pairs of this CTA
  const int n_stage_it = my_ts > 0 ? my_ts * kGtSub : 0;
  const bool unit = prm.ckpt_unit != 0;
  const bool gen_ze = prm.ckpt_zeta == 0;      // zeta is not in the checkpoint: regenerate it from Philox (see RolloutParams)
  // per-path cotangents of the 4 samples of quad j of stage iteration `it` (forward-written checkpoint, unit cotangents)
  auto unit_w = [&](int it, int j, float (&w)[4]) {
    const int ts = (int)blockIdx.x + (it / kGtSub) * (int)gridDim.x, sub = it % kGtSub;
    const int k0 = (prm.tile0 + ts / prm.N) * kCkP + sub * kGtS + 4 * j;
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (k0 + i < prm.K_local) ? __ldg(prm.wY + k0 + i) : 0.f;
  };
  // forward-written checkpoint: scale the zeta rows by the per-path cotangent and drop (zero) every row of a path whose
  // cotangent is zero, so that a diverged trajectory the loss has discarded cannot poison the sums.  Threads 0..575
  // (delta + lo warps); thread t owns chunk position t & 7 of rows (t >> 3) + 72 i: its sample quad is fixed.
  constexpr int kFix = kGtWorkers + 64;
  auto fixup = [&](uint8_t* tH, int it, int t) {
    const int pos = t & 7, r0 = t >> 3;
    const int j = pos ^ (r0 & 7);
    float w[4];
    unit_w(it, j, w);
    const int r_end = gen_ze ? tg.act_rows : tg.cols;          // regenerated zeta rows already carry the cotangent
    for (int r = r0; r < r_end; r += kFix / 8) {
      float4* p = reinterpret_cast<float4*>(tH + (uint32_t)r * 128u + (uint32_t)(pos << 4));
      float4 v = *p;
      const bool ze = r >= tg.r_ze;
      v.x = w[0] != 0.f ? (ze ? v.x * w[0] : v.x) : 0.f; v.y = w[1] != 0.f ? (ze ? v.y * w[1] : v.y) : 0.f;
      v.z = w[2] != 0.f ? (ze ? v.z * w[2] : v.z) : 0.f; v.w = w[3] != 0.f ? (ze ? v.w * w[3] : v.w) : 0.f;
      *p = v;
    }
    tc::fence_proxy_async();
    gt_named_bar(1, kFix);
  };

  if (warp < kGtUtil) {
    // =============================================================== hidden cotangents (512 threads)
    // warp = 4 hidden columns hc0..hc0+3 of [h1 (32) | h2 (32)]; lane = 4 samples (quad hj) x quarter hk of the reduction range
    const int hj = lane & 7, hk = lane >> 3;
    const int hc0 = 4 * warp;
    const bool is_h2 = hc0 >= 32;
    const int seg_n = g.dims[is_h2 ? 2 : 1];
    const int h_row = tg.s0 + hc0;                              // tile row of the hidden activation h[hc0]
    // k4 = hk, hk + 4, ...: the cotangent rows row0 + 4 k4 + e of this lane keep (row & 7) = 4 (hk & 1) + e (row0 % 8 == 0),
    // so the swizzled chunk of the lane's sample quad is a per-lane constant for each e
    uint32_t ch[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) ch[e] = (uint32_t)(4 * hk + e) * 128u + (uint32_t)((hj ^ (4 * (hk & 1) + e)) << 4);
    PhaseTimer pt_;      // debug (CTA 0, thread 0): [0] wait for the TMA, [1] zeta . W2' (+ delta_2), [2] barrier + delta_2 . W1' + delta_1, [3] fence + arrive
    pt_.start(prm.prof, tid);
    for (int it = 0; it < n_stage_it; ++it) {
      const int s = it & 1;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      uint8_t* tH = smem + tg.o_hi[s];
      uint8_t* tL = smem + tg.o_lo[s];
      if (gen_ze) {
        // zeta rows of this stage from Philox, exactly as the rollout drew them: thread = (sample, column group); value
        // wY[path] * (sqrt(dt) xi), zero for padding paths and dropped trajectories.  The stage buffer must be free (the
        // tensor core is done with its previous use); the TMA load of the activation rows is in flight meanwhile.
        tc::mbar_wait(&bar_empty[s], par ^ 1u);
        const int ts = (int)blockIdx.x + (it / kGtSub) * (int)gridDim.x, sub = it % kGtSub;
        const int kl = (prm.tile0 + ts / prm.N) * kCkP + sub * kGtS + lane;          // local path index of this lane's sample
        const unsigned nstep = (unsigned)(ts % prm.N);
        const float wk = (kl < prm.K_local) ? __ldg(prm.wY + kl) : 0.f;
        const float sqdt = sqrtf(prm.dt);
        const int ngroups = (prm.d + 3) >> 2;
        for (int jb = warp; jb < ngroups; jb += kGtWorkers / 32) {
          const float4 e4 = philox_normal4((unsigned)(prm.k_offset + kl), nstep, (unsigned)jb, prm.offset, prm.seed);
          const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = 4 * jb + i;
            const float v = (c < prm.d && wk != 0.f) ? wk * (sqdt * ev[i]) : 0.f;
            const uint32_t off = (uint32_t)(tg.r_ze + c) * 128u + (uint32_t)((((lane >> 2) ^ ((tg.r_ze + c) & 7)) << 4) + ((lane & 3) << 2));
            *reinterpret_cast<float*>(tH + off) = v;
            *reinterpret_cast<float*>(tL + off) = lo1(v);
          }
        }
        gt_named_bar(3, kGtWorkers);
      } else {
        tc::mbar_wait(&bar_full[s], par);
      }
      pt_.mark(0);
      if (unit && !gen_ze) fixup(tH, it, tid);
      float acc[4][4];                      // [hidden column][sample]
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
      // cotangent rows row0 + 4 k4 + e, weights W[hc0 + c][4 k4 + e]
      auto accumulate = [&](const float* w, int nng, int row0, int nk4) {
        const uint8_t* zb = tH + (uint32_t)row0 * 128u;
        const float* wb = w + (size_t)warp * nng * 16;
#pragma unroll 2
        for (int k4 = hk; k4 < nk4; k4 += 4) {
          float4 z[4], wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) z[e] = *reinterpret_cast<const float4*>(zb + (uint32_t)(k4 - hk) * 512u + ch[e]);
#pragma unroll
          for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(wb + k4 * 16 + c * 4);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float we[4] = {wv[c].x, wv[c].y, wv[c].z, wv[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[c][0] = fmaf(z[e].x, we[e], acc[c][0]); acc[c][1] = fmaf(z[e].y, we[e], acc[c][1]);
              acc[c][2] = fmaf(z[e].z, we[e], acc[c][2]); acc[c][3] = fmaf(z[e].w, we[e], acc[c][3]);
            }
          }
        }
      };
      auto finish = [&](int row_dst) {       // combine the four quarters; lane hk finishes column hk: act', zero the pads, raw -> tH, lo -> tL
        float sel[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v = acc[c][i];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (c == hk) sel[i] = v;
          }
        const int c = hk;
        const float4 h = *reinterpret_cast<const float4*>(tH + gt_swz(h_row + c, hj));
        const float hv[4] = {h.x, h.y, h.z, h.w};
        const bool live = (hc0 & 31) + c < seg_n;
        float v[4];
        // act': relu(.)^2 -> 2 relu(pre) = 2 sqrt(h) (sqrt.approx: 1 ulp, far inside the 1e-5 parity bar); tanh -> 1 - h^2
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = live ? sel[i] * (tg.dense ? 2.0f * gt_sqrt_approx(hv[i]) : (1.0f - hv[i] * hv[i])) : 0.f;
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        const uint32_t off = gt_swz(row_dst + (hc0 & 31) + c, hj);
        *reinterpret_cast<float4*>(tH + off) = o;
        *reinterpret_cast<float4*>(tL + off) = lo4(o);
      };
      if (is_h2 || tg.dense) accumulate(sW + tg.o_w2, tg.w2_nng, tg.r_ze, g.layer[2].nng);
      if (gen_ze) {                           // the hidden activations (act') arrive by TMA: needed from here on
        tc::mbar_wait(&bar_full[s], par);
        if (unit) fixup(tH, it, tid);
      }
      if (is_h2) finish(tg.r_d2);
      pt_.mark(1);
      gt_named_bar(2, kGtWorkers);
      if (!is_h2) {
        accumulate(sW + tg.o_w1, tg.w1_nng, tg.r_d2, g.layer[1].nng);
        finish(tg.r_d1);
      }
      pt_.mark(2);
      tc::fence_proxy_async();
      tc::mbar_arrive(&bar_dl[s]);
      pt_.mark(3);
    }
  } else {
    // =============================================================== utility warps 16..19
    const int qtr = warp & 3;                                  // tensor-memory lane quarter of this warp (flush)
    // raw accumulators -> this CTA's partial, [m tile][cotangent column][lane = checkpoint column] (coalesced);
    // reduce_grad_tc_kernel sums the partials in fp64 and scatters them to theta.  The tensor core does not round its
    // FP32 accumulation to nearest: the error of a sum kept in tensor memory grows linearly with the number of MMAs that
    // went into it (measured: 1.1e-5 relative after 1 200 against 1.3e-6 after 48), so the accumulators are added to the
    // partial (round-to-nearest FP32 adds, L2 resident, one writer per address) every flush_stages stages and started over.
    auto flush = [&](uint32_t fpar) {
      tc::mbar_wait(bar_acc_full, fpar);
      tc::fence_after_sync();
      float* gp = prm.grad_partial + (size_t)blockIdx.x * (2 * 128 * tg.nB);
      for (int mt = 0; mt < tg.n_mt; ++mt) {
        if (128 * mt + 32 * qtr >= tg.act_rows) continue;      // lanes past the last activation column hold garbage rows
        for (int c0 = 0; c0 < tg.nB; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tbase + (((uint32_t)(32 * qtr)) << 16) + (uint32_t)(mt * tg.nB + c0), v);
          tc::wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(gp + (size_t)(mt * tg.nB + c0 + i) * 128 + 32 * qtr + lane, v[i]);   // RED
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_acc_empty);
    };
    // debug (CTA 0): MMA lane [4] wait accumulators free + lo tile, [5] wait delta rows, [6] issue + commit, [7] flush;
    // lo thread [8] wait for the TMA, [9] lo pass, [10] flush; TMA lane [11] wait for a free stage, [12] issue, [13] flush
    PhaseTimer pt_;
    pt_.start(prm.prof, (lane == 0 && warp <= kGtUtil + 2) ? 0 : 1);
    const int pb = warp == kGtUtil + 1 ? 4 : warp == kGtUtil + 2 ? 8 : 11;
    uint32_t n_flush = 0;
    for (int it = 0; it < n_stage_it; ++it) {
      const int s = it & 1;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      uint8_t* tH = smem + tg.o_hi[s];
      uint8_t* tL = smem + tg.o_lo[s];
      const bool flush_now = ((it + 1) % flush_stages == 0) || it == n_stage_it - 1;
      if (warp == kGtUtil) {
        // ----------------------------------------------------------- TMA producer
        if (lane == 0) {
          tc::mbar_wait(&bar_empty[s], par ^ 1u);                  // first use of a stage passes immediately
          pt_.mark(11);
          const int ts = (int)blockIdx.x + (it / kGtSub) * (int)gridDim.x, sub = it % kGtSub;
          if (gen_ze) {                       // activation rows only: one box of act_rows rows (the map's box)
            tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)tg.act_rows * 128u);
            tc::tma_load_3d(tH, &tmap, &bar_full[s], sub * kGtS, 0, ts);
          } else {
            tc::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)tg.cols * 128u);
            tc::tma_load_3d(tH, &tmap, &bar_full[s], sub * kGtS, 0, ts);
            tc::tma_load_3d(tH + (uint32_t)tg.box_rows * 128u, &tmap, &bar_full[s], sub * kGtS, tg.box_rows, ts);
          }
          pt_.mark(12);
        }
        __syncwarp();
      } else if (warp == kGtUtil + 1) {
        // ----------------------------------------------------------- MMA issuer
        if (lane == 0) {
          const uint32_t sH = tc::smem_u32(tH), sL = tc::smem_u32(tL);
          const bool first = (it % flush_stages) == 0;             // accumulators start over after a flush
          // D[m tile][:, dcol .. dcol + n) (+)= act' . cot rows [row0, row0 + n) over the 32 samples of the stage, three passes
          auto issue = [&](int row0, int n, int dcol) {
            const uint32_t id = tc::idesc_tf32(128, n);
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = (pass == 0) ? sL : sH;
              const uint32_t b = ((pass == 1) ? sL : sH) + (uint32_t)row0 * 128u;
              for (int mt = 0; mt < tg.n_mt; ++mt)
                for (int ks = 0; ks < kGtS / 8; ++ks) {
                  const uint64_t ad = tc::smem_desc_sw128(a + (uint32_t)mt * 16384u + (uint32_t)ks * 32u);
                  const uint64_t bd = tc::smem_desc_sw128(b + (uint32_t)ks * 32u);
                  tc::mma_tf32_ss(tbase + (uint32_t)(mt * tg.nB + dcol), ad, bd, id, !first || pass > 0 || ks > 0);
                }
            }
          };
          if (first && n_flush > 0) { tc::mbar_wait(bar_acc_empty, (n_flush - 1u) & 1u); }
          tc::mbar_wait(&bar_lo[s], par);
          pt_.mark(4);
          tc::mbar_wait(&bar_dl[s], par);
          pt_.mark(5);
          tc::fence_after_sync();
          issue(tg.r_ze, tg.nB, 0);          // one group of N = nB: a K = 8 MMA does not get cheaper in proportion to N
          tc::mma_commit(&bar_empty[s]);
          if (flush_now) tc::mma_commit(bar_acc_full);
          pt_.mark(6);
        }
        __syncwarp();
      } else {
        // ----------------------------------------------------------- lo tile of the loaded rows (64 threads)
        tc::mbar_wait(&bar_full[s], par);
        pt_.mark(8);
        if (unit) fixup(tH, it, tid - 64);                         // threads 576..639 -> 512..575
        const int t = tid - (kGtUtil + 2) * 32;
        const float4* src = reinterpret_cast<const float4*>(tH);
        float4* dst = reinterpret_cast<float4*>(tL);
        const int nq = (gen_ze ? tg.act_rows : tg.cols) * 8;       // regenerated zeta rows come with their lo values
        int q = t;
        for (; q + 7 * 64 < nq; q += 8 * 64) {                     // 8 loads in flight per thread: the shared-memory latency
          float4 v[8];                                             // under the tensor core's operand traffic is > 100 cycles
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = src[q + 64 * u];
#pragma unroll
          for (int u = 0; u < 8; ++u) dst[q + 64 * u] = lo4(v[u]);
        }
        for (; q < nq; q += 64) dst[q] = lo4(src[q]);
        tc::fence_proxy_async();
        tc::mbar_arrive(&bar_lo[s]);
        pt_.mark(9);
      }
      if (flush_now) { flush(n_flush & 1u); ++n_flush; pt_.mark(pb + (warp == kGtUtil + 1 ? 3 : 2)); }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// partial[cta][m tile][cotangent column c][lane m] -> grad_theta.  Lane m of M tile mt is checkpoint column 128 mt + m
// (segments padded to s0 / 32 / 32), column c is cotangent column c of [zeta | delta_2 | delta_1]; every parameter is
// produced by exactly one (column, row) pair.  Fixed summation order, fp64.
static __global__ void reduce_grad_tc_kernel(const NetGeom g, const GradTcGeom tg, const float* __restrict__ partial, int nparts,
                                             float* __restrict__ out) {
  const int per = 2 * 128 * tg.nB;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
    const int lane = q & 127, mc = q >> 7, mt = mc / tg.nB, c = mc - mt * tg.nB;
    if (mt >= tg.n_mt) continue;
    const int m = 128 * mt + lane;
    int col = -1;
    if (m < tg.s0) { if (m < g.seg_len[0]) col = m; }
    else if (m < tg.s0 + 32) { if (m - tg.s0 < g.seg_len[1]) col = g.seg_off[1] + (m - tg.s0); }
    else if (m < tg.s0 + 64) { if (m - tg.s0 - 32 < g.seg_len[2]) col = g.seg_off[2] + (m - tg.s0 - 32); }
    if (col < 0) continue;
    int l, n;
    if (c < tg.s0) { l = 2; n = c; }
    else if (c < tg.nE) continue;
    else if (c < tg.nE + 32) { l = 1; n = c - tg.nE; }
    else { l = 0; n = c - tg.nE - 32; }
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

// ---- host: tensor map of a checkpoint buffer [n_ts][cols][128 paths] fp32, box = 32 paths x box_rows columns x 1,
//      128-byte swizzle.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmap_encode tmap_encode_fn() {
  static PFN_tmap_encode fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmap_encode>(p);
  }
  return fn;
}
// returns 0 on success
inline int grad_tc_tensor_map(const GradTcGeom& tg, const float* ckpt, long long n_ts, CUtensorMap* out, int box_rows = 0, int cols = 0) {
  if (box_rows <= 0) box_rows = tg.box_rows;
  if (cols <= 0) cols = tg.cols;              // columns per (slot, step) in memory (RolloutParams::ckpt_cols)
  PFN_tmap_encode enc = tmap_encode_fn();
  if (!enc) return -1;
  const cuuint64_t gdim[3] = {(cuuint64_t)kCkP, (cuuint64_t)cols, (cuuint64_t)n_ts};
  const cuuint64_t gstr[2] = {(cuuint64_t)kCkP * 4u, (cuuint64_t)kCkP * 4u * (cuuint64_t)cols};
  const cuuint32_t box[3] = {(cuuint32_t)kGtS, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ckpt), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace pspde
#endif  // !PSPDE_EMULATE
