// grad_tc_kernels.cuh -- gradient accumulation from the checkpoint rows with the WEIGHT GRADIENT on the 5th-generation
// tensor cores (tcgen05 + tensor memory).  Same contract as grad_kernel (grad_kernels.cuh): a pure streaming kernel
// over (trajectory, step) samples, dtheta = sum_samples J_theta Z(a0)' zeta.
//
// The reduction dimension K of the weight gradient
//     D[act col][cot col] += sum_samples act[sample][act col] * cot[sample][cot col]
// is the SAMPLE index, M = activation column, N = cotangent column.  A half tile (64 samples) of checkpoint rows is
// copied into shared memory with a 4 x 4 register transpose so that four consecutive samples of one column form a
// float4: element (row r, sample 4 j + i) at j * LBO + r * 16 + i * 4 bytes, the canonical no-swizzle K-major operand
// layout of tcgen05.mma (8-row x 16-byte core matrices, SBO = 128 B).  All columns -- [a0 | h1 | h2] and
// [zeta | delta_2 | delta_1] -- are rows of ONE such matrix, so the A operand (two M tiles: rows 0..127, 128..255) and
// the B operand (rows from zeta on: zeta padded to a multiple of 16 | delta_2 | delta_1, N = 176 at the C2 shape) are
// windows of the same tile; the zeta columns are issued right after the copy (beside the hidden-cotangent FMA code),
// the delta columns when they exist.  (MN-major descriptors,
// which would accept the checkpoint layout without the transpose, returned zeros for kind::tf32 without swizzle.)
// FP32 equivalence as in tc_sm100.cuh: the tile holds the raw FP32 values, which the tensor core truncates to TF32
// (hi = trunc(x)), a second tile holds lo = x - trunc(x); three passes lo*hi + hi*lo + hi*hi into the FP32
// accumulators in tensor memory (2 M tiles x N columns), which stay resident for the whole launch and are added to
// the CTA's gradient partial once at the end.
//
// The hidden cotangents delta_2, delta_1 (6 900 of the 29 960 MACs per sample at the C2 shape) are formed by FP32
// FMA code on 4 samples x 4 hidden columns per thread that reads zeta / h straight from the operand tile (one
// LDS.128 = one column of a sample quad) and appends the result to it.
#pragma once
#if !defined(PSPDE_EMULATE)
#include "grad_kernels.cuh"
#include "tc_sm100.cuh"

namespace pspde {

constexpr int kGtS = 64;                 // samples per work item (half a checkpoint tile) = K of one MMA group
constexpr int kGtThreads = 512;
constexpr int kGtQ = kGtS / 4;           // sample quads per item
constexpr int kGtFlushItems = 4;         // accumulator flush period (items of 64 samples = 48 MMAs per accumulator)

struct GradTcGeom {
  int s04;             // column groups of a0 in the checkpoint (s0 / 4)
  int act_groups;      // s04 + 16: [a0 | h1 | h2]
  int ze_groups;       // s04: zeta (d used, the rest zero)
  int g_ze, g_d2, g_d1, n_groups;        // first group of zeta / delta_2 / delta_1 in the tile; total incl. zero pad
  int nB;              // N of the weight-gradient product: zeta (padded to nE) | delta_2 (32) | delta_1 (32)
  int nE;              // zeta columns padded to a multiple of 16: their MMAs are issued before the hidden cotangents exist
  int dense;
  int n_mt;            // M tiles of 128 activation columns (1 when [a0 | h1 | h2] has at most 128 columns)
  uint32_t lbo;        // bytes between sample quads of the operand tile: (4 n_groups + 1) * 16
  // compact weights for the hidden cotangents, k4-blocked: W2c rows = [h1 (32) | h2 (32)] columns (zero rows where
  // the layer does not read the column), W1c rows = h1 (32) columns
  int w2_nng, w1_nng, o_w2, o_w1;
  uint32_t o_hi, o_lo, o_w, total;       // shared-memory byte offsets
};

inline bool grad_tc_geom(const NetGeom& g, int d, int s0, GradTcGeom& t) {
  if (g.L != 3 || g.time_mode == TIME_NONE || g.seg_len[1] > 32 || g.seg_len[2] > 32 || (s0 & 7) || s0 < g.seg_len[0]) return false;
  t.dense = g.kind == NET_DENSENET ? 1 : 0;
  t.s04 = s0 >> 2;
  t.act_groups = t.s04 + 16;
  t.ze_groups = t.s04;
  t.nE = ((4 * t.ze_groups + 15) / 16) * 16;
  t.nB = t.nE + 64;
  t.g_ze = t.act_groups; t.g_d2 = t.g_ze + t.nE / 4; t.g_d1 = t.g_d2 + 8;
  t.n_groups = t.g_ze + t.nB / 4;
  t.n_mt = t.act_groups > 32 ? 2 : 1;
  if (t.n_groups < 32 * t.n_mt) t.n_groups = 32 * t.n_mt;     // M tile mt reads rows [128 mt, 128 mt + 128)
  if (t.act_groups > 64 || t.nB > 256 || 2 * t.nB > 512) return false;
  (void)d;
  t.lbo = (uint32_t)(4 * t.n_groups + 1) * 16u;    // odd number of 16-byte rows: 8 consecutive quads hit 8 distinct bank groups
  t.w2_nng = g.layer[2].nng; t.w1_nng = g.layer[1].nng;
  uint32_t o = 0;
  t.o_hi = o; o += (uint32_t)kGtQ * t.lbo;
  t.o_lo = o; o += (uint32_t)kGtQ * t.lbo;
  t.o_w = o;
  t.o_w2 = 0; t.o_w1 = 64 * t.w2_nng * 4;
  o += (uint32_t)(64 * t.w2_nng * 4 + 32 * t.w1_nng * 4) * 4u;
  t.total = o;
  return t.total <= 227u * 1024u;
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// lo = x - trunc(x), rounded to TF32 so that the tensor core's own truncation of it is exact
__device__ __forceinline__ float lo1(float x) { return tc::tf32_hi(x - trunc_tf32(x)); }
__device__ __forceinline__ float4 lo4(const float4& v) { return make_float4(lo1(v.x), lo1(v.y), lo1(v.z), lo1(v.w)); }
// one checkpoint float4 of a sample whose cotangent w is applied here (RolloutParams::ckpt_unit): zeta columns are scaled,
// activation columns kept; w == 0 drops the sample (selects, no branches)
__device__ __forceinline__ float4 scale_row(const float4& v, float w, bool is_ze) {
  const float f = is_ze ? w : 1.0f;
  const bool keep = w != 0.f;
  return make_float4(keep ? v.x * f : 0.f, keep ? v.y * f : 0.f, keep ? v.z * f : 0.f, keep ? v.w * f : 0.f);
}

__global__ void __launch_bounds__(kGtThreads, 1) grad_tc_kernel(const RolloutParams prm, const GradTcGeom tg, const int n_items,
                                                                const int flush_items) {
  extern __shared__ float4 smem4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const NetGeom& g = prm.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* tH = smem + tg.o_hi;
  uint8_t* tL = smem + tg.o_lo;
  float* sW = reinterpret_cast<float*>(smem + tg.o_w);
  const uint32_t lbo = tg.lbo;
  // element (row r = column of [a0 | h1 | h2 | zeta | delta_2 | delta_1], sample 4 j + i) sits at j * lbo + r * 16 + i * 4:
  // the canonical no-swizzle K-major operand layout with K = sample (8 rows x 16 B core matrices, SBO = 128 B, LBO = lbo)
  auto quad = [&](uint8_t* t, int r, int j) -> float4& { return *reinterpret_cast<float4*>(t + (uint32_t)j * lbo + (uint32_t)r * 16u); };

  // ---- one-time setup
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar_mma, 1); tc::mbar_fence_init(); }
  for (uint32_t q = tid; q < 2u * kGtQ * lbo / 16u; q += kGtThreads) reinterpret_cast<float4*>(tH)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  const uint32_t sH = tc::smem_u32(tH), sL = tc::smem_u32(tL);

  const float4* ck = reinterpret_cast<const float4*>(prm.ckpt);
  const int src_groups = tg.act_groups + tg.ze_groups;       // the checkpoint row: [a0 | h1 | h2 | zeta]
  const bool unit = prm.ckpt_unit != 0;
  // hidden-cotangent mapping: warp = 4 hidden columns hc0 .. hc0 + 3 of [h1 (32) | h2 (32)], lane = (sample quad, half of
  // the reduction range); the two halves are combined with one xor-shuffle
  const int hc0 = 4 * warp, hj = lane & 15, hk = lane >> 4;
  const bool is_h2 = hc0 >= 32;
  const int seg_n = g.dims[is_h2 ? 2 : 1];
  const int h_row = 4 * tg.s04 + hc0;                         // tile row of the hidden activation h[hc0]
  // D[m tile][:, n0 .. n0 + n) += act' . cot[:, n0 .. n0 + n) over the 64 samples of the tile, three passes (thread 0)
  auto issue = [&](int n0, int n, bool overwrite) {
    const uint32_t id = tc::idesc_tf32(128, n);
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a = (pass == 0) ? sL : sH;
      const uint32_t b = ((pass == 1) ? sL : sH) + (uint32_t)(4 * tg.g_ze + n0) * 16u;
      for (int mt = 0; mt < tg.n_mt; ++mt)
        for (int ks = 0; ks < kGtS / 8; ++ks) {
          const uint64_t ad = tc::smem_desc(a + (uint32_t)mt * 2048u + (uint32_t)ks * 2u * lbo, lbo, 128u);
          const uint64_t bd = tc::smem_desc(b + (uint32_t)ks * 2u * lbo, lbo, 128u);
          mma_tf32_ss(tbase + (uint32_t)(mt * tg.nB + n0), ad, bd, id, !overwrite || pass > 0 || ks > 0);
        }
    }
  };
  // raw accumulators -> this CTA's partial, [m tile][cotangent column][lane = checkpoint column] (coalesced);
  // reduce_grad_tc_kernel sums the partials in fp64 and scatters them to theta.  A warp reads the lane quarter 32 (warp % 4).
  // Called every flush_items items (see the note at the call site).
  auto flush = [&]() {
    float* gp = prm.grad_partial + (size_t)blockIdx.x * (2 * 128 * tg.nB);
    const int qtr = warp & 3, cpart = warp >> 2;          // 4 warps per lane quarter split the columns
    for (int mt = 0; mt < tg.n_mt; ++mt) {
      if (128 * mt + 32 * qtr >= 4 * tg.act_groups) continue;      // lanes past the last activation column hold garbage rows
      for (int c0 = 8 * cpart; c0 < tg.nB; c0 += 32) {
        float v[8];
        tc::tmem_ld8(tbase + (((uint32_t)(32 * qtr)) << 16) + (uint32_t)(mt * tg.nB + c0), v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(gp + (size_t)(mt * tg.nB + c0 + i) * 128 + 32 * qtr + lane, v[i]);   // RED: no round trip; one writer per address
      }
    }
  };
  uint32_t ph = 0;
  bool first = true, pending = false;
  int since_flush = 0;
  PhaseTimer pt_;          // debug: [0] wait for the tensor core, [1] copy + transpose, [2] hidden cotangents, [3] MMA issue, [4] zeta . W2', [5] delta_2 + barrier, [6] delta_2 . W1' + delta_1, [2] fences + barrier
  pt_.start(prm.prof, tid == 32 ? 0 : 1);

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int ts = item / 2, half = item - ts * 2;
    const float4* src = ck + (size_t)ts * prm.ckpt_c4 * kCkP + half * kGtS;
    // ---- (1) copy [a0 | h1 | h2 | zeta] with a 4 x 4 register transpose (sample-major float4 -> sample quads per column);
    //      all global loads of the thread are in flight before the first one is used
    {
      constexpr int MAXI = 3;
      float4 v[MAXI][4];
      const int nq = src_groups * kGtQ;
      // checkpoint written by the forward pass (unit cotangents): the per-path cotangent dL/dY_N of the thread's 4 samples
      // (its sample quad is tid & 15 in every iteration below) scales the zeta rows; a path with zero cotangent is dropped
      // entirely, so that a diverged trajectory (non-finite rows) with zero weight cannot poison the sums.
      float wq[4] = {1.f, 1.f, 1.f, 1.f};
      if (unit) {
        const int k0 = (prm.tile0 + ts / prm.N) * kCkP + half * kGtS + 4 * (tid & (kGtQ - 1));
#pragma unroll
        for (int i = 0; i < 4; ++i) wq[i] = (k0 + i < prm.K_local) ? __ldg(prm.wY + k0 + i) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < MAXI; ++it) {
        const int q = tid + it * kGtThreads;
        if (q < nq) {
          const float4* sp = src + (size_t)(q >> 4) * kCkP + 4 * (q & (kGtQ - 1));
          v[it][0] = __ldg(sp); v[it][1] = __ldg(sp + 1); v[it][2] = __ldg(sp + 2); v[it][3] = __ldg(sp + 3);
          if (unit) {
            const bool is_ze = (q >> 4) >= tg.act_groups;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[it][i] = scale_row(v[it][i], wq[i], is_ze);
          }
        }
      }
      // the tensor core must be done with the tile of the previous item before it is overwritten (the loads above are
      // in flight meanwhile)
      if (pending) { tc::mbar_wait(&bar_mma, ph); ph ^= 1u; pending = false; tc::fence_after_sync(); }
      // The tensor core does not round its FP32 accumulation to nearest: the error of a sum kept in tensor memory grows
      // linearly with the number of MMAs that went into it (measured: 1.1e-5 relative after 25 items against 1.3e-6
      // after one).  Every flush_items items the accumulators are therefore added to the CTA's partial (round-to-nearest
      // FP32 adds, L2 resident) and started over -- here, where the pipeline has to wait for the tensor core anyway.
      if (since_flush >= flush_items) { flush(); first = true; since_flush = 0; }
      pt_.mark(0);
#pragma unroll
      for (int it = 0; it < MAXI; ++it) {
        const int q = tid + it * kGtThreads;
        if (q < nq) {
          const int gi = q >> 4, j = q & (kGtQ - 1);
          const float4 c0 = make_float4(v[it][0].x, v[it][1].x, v[it][2].x, v[it][3].x), c1 = make_float4(v[it][0].y, v[it][1].y, v[it][2].y, v[it][3].y);
          const float4 c2 = make_float4(v[it][0].z, v[it][1].z, v[it][2].z, v[it][3].z), c3 = make_float4(v[it][0].w, v[it][1].w, v[it][2].w, v[it][3].w);
          quad(tH, 4 * gi, j) = c0; quad(tH, 4 * gi + 1, j) = c1; quad(tH, 4 * gi + 2, j) = c2; quad(tH, 4 * gi + 3, j) = c3;
          quad(tL, 4 * gi, j) = lo4(c0); quad(tL, 4 * gi + 1, j) = lo4(c1); quad(tL, 4 * gi + 2, j) = lo4(c2); quad(tL, 4 * gi + 3, j) = lo4(c3);
        }
      }
      for (int q = tid + MAXI * kGtThreads; q < nq; q += kGtThreads) {                  // (wider inputs than the C2 shape)
        const int gi = q >> 4, j = q & (kGtQ - 1);
        const float4* sp = src + (size_t)gi * kCkP + 4 * j;
        float4 v0 = __ldg(sp), v1 = __ldg(sp + 1), v2 = __ldg(sp + 2), v3 = __ldg(sp + 3);
        if (unit) {
          const bool is_ze = gi >= tg.act_groups;
          v0 = scale_row(v0, wq[0], is_ze); v1 = scale_row(v1, wq[1], is_ze); v2 = scale_row(v2, wq[2], is_ze); v3 = scale_row(v3, wq[3], is_ze);
        }
        const float4 c0 = make_float4(v0.x, v1.x, v2.x, v3.x), c1 = make_float4(v0.y, v1.y, v2.y, v3.y);
        const float4 c2 = make_float4(v0.z, v1.z, v2.z, v3.z), c3 = make_float4(v0.w, v1.w, v2.w, v3.w);
        quad(tH, 4 * gi, j) = c0; quad(tH, 4 * gi + 1, j) = c1; quad(tH, 4 * gi + 2, j) = c2; quad(tH, 4 * gi + 3, j) = c3;
        quad(tL, 4 * gi, j) = lo4(c0); quad(tL, 4 * gi + 1, j) = lo4(c1); quad(tL, 4 * gi + 2, j) = lo4(c2); quad(tL, 4 * gi + 3, j) = lo4(c3);
      }
    }
    // the zeta part of the weight gradient does not need the hidden cotangents: its MMAs run beside them
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) { tc::fence_after_sync(); issue(0, tg.nE, first); }
    if (item + (int)gridDim.x < n_items) {                    // next item -> L2 while the hidden cotangents are formed
      const int nitem = item + gridDim.x, nts = nitem / 2, nhalf = nitem - nts * 2;
      for (int q = tid; q < src_groups * 8; q += kGtThreads) {                           // 8 lines of 128 B per column group
        const char* np_ = reinterpret_cast<const char*>(ck + (size_t)nts * prm.ckpt_c4 * kCkP + nhalf * kGtS + (size_t)(q >> 3) * kCkP) + (q & 7) * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(np_));
      }
    }
    pt_.mark(1);
    // ---- (2) hidden cotangents for 4 samples x 4 hidden columns per thread:
    //      dh[s][c] = sum_n zeta[s][n] W2[c][n]   (+ sum_n delta_2[s][n] W1[c][n] for the h1 columns)
    float acc[4][4];                      // [hidden column][sample]
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
    auto accumulate = [&](const float* w, int nng, int row0, int nk4) {     // cotangent rows row0 + 4 k4 + e, weights W[hc0 + c][4 k4 + e]
      for (int k4 = hk; k4 < nk4; k4 += 2) {
        float4 z[4], wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) z[e] = quad(tH, row0 + 4 * k4 + e, hj);
#pragma unroll
        for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const float4*>(w + ((hc0 >> 2) * nng + k4) * 16 + c * 4);
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
    auto finish = [&](int row_dst) {       // combine the halves, act', zero the pads, raw -> tH, lo -> tL
      float sel[2][4];                     // half hk of the reduction lanes finishes columns 2 hk, 2 hk + 1 (branch-free selects)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v = acc[c][i] + __shfl_xor_sync(0xffffffffu, acc[c][i], 16);
          if ((c >> 1) == hk) sel[c & 1][i] = v;
        }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * hk + cc;
        const float4 h = quad(tH, h_row + c, hj);
        const float hv[4] = {h.x, h.y, h.z, h.w};
        const bool live = (hc0 & 31) + c < seg_n;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = live ? sel[cc][i] * (tg.dense ? 2.0f * sqrtf(hv[i]) : (1.0f - hv[i] * hv[i])) : 0.f;
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        quad(tH, row_dst + (hc0 & 31) + c, hj) = o;
        quad(tL, row_dst + (hc0 & 31) + c, hj) = lo4(o);
      }
    };
    if (is_h2 || tg.dense) accumulate(sW + tg.o_w2, tg.w2_nng, 4 * tg.g_ze, g.layer[2].nng);
    pt_.mark(4);
    if (is_h2) finish(4 * tg.g_d2);
    __syncthreads();
    pt_.mark(5);
    if (!is_h2) {
      accumulate(sW + tg.o_w1, tg.w1_nng, 4 * tg.g_d2, g.layer[1].nng);
      finish(4 * tg.g_d1);
    }
    pt_.mark(6);
    // ---- (3) weight gradient: D[m tile][act col][cot col] += sum_samples, three passes, one thread issues
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    pt_.mark(2);
    if (tid == 0) {
      tc::fence_after_sync();
      issue(tg.nE, 64, first);
      tc::mma_commit(&bar_mma);
    }
    first = false;
    pending = true;
    pt_.mark(3);
    ++since_flush;
  }
  if (pending) { tc::mbar_wait(&bar_mma, ph); tc::fence_after_sync(); }

  if (!first) flush();
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
    if (m < 4 * tg.s04) { if (m < g.seg_len[0]) col = m; }
    else if (m < 4 * tg.s04 + 32) { if (m - 4 * tg.s04 < g.seg_len[1]) col = g.seg_off[1] + (m - 4 * tg.s04); }
    else if (m < 4 * tg.s04 + 64) { if (m - 4 * tg.s04 - 32 < g.seg_len[2]) col = g.seg_off[2] + (m - 4 * tg.s04 - 32); }
    if (col < 0) continue;
    int l, n;
    if (c < 4 * tg.ze_groups) { l = 2; n = c; }
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

}  // namespace pspde
#endif  // !PSPDE_EMULATE
