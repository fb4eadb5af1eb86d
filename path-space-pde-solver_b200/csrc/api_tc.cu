// api_tc.cu -- tensor-core (tcgen05 / tensor memory) building block and its self test.
#include "api_common.h"
#include "tc_sm100.cuh"
#include "grad_tc_kernels.cuh"

#if !defined(PSPDE_EMULATE)
namespace pspde {

// D[128 x N] = A[128 x K] . B[K x N] through the 3xTF32 path of tc_sm100.cuh.  One CTA, 160 threads: warps 0-3 own
// one row each (A -> tensor memory, D -> global), warp 4 lane 0 issues the MMAs.  K % 8 == 0, N % 16 == 0,
// 2K + N <= 512.
__global__ void __launch_bounds__(160, 1) tc_selftest_kernel(int K, int N, int variant, const float* __restrict__ A,
                                                             const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ float4 smem4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tile_bytes = tc::b_tile_bytes(K, N);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + tile_bytes;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  // B tile: hi / lo copies in the K-major core-matrix layout
  // variant bit 2 (experiment): B in the MN-major 128-byte-swizzled canonical layout (N == 32: one 128 B row of N per k,
  // 8 k rows = one 1 KB atom, 16-byte chunk index xor-ed with the k row inside the atom); tiles 1 KB aligned
  const uint32_t align_pad = (variant & 4) ? ((1024u - (tc::smem_u32(smem) & 1023u)) & 1023u) : 0u;
  if (variant & 4) { b_hi = smem + align_pad; b_lo = b_hi + ((tile_bytes + 1023u) & ~1023u); }
  for (int q = tid; q < K * N; q += blockDim.x) {
    const int k = q / N, n = q - k * N;
    float hi, lo;
    tc::tf32_split_rn(B[q], hi, lo);
    // bit 5: the SWIZZLE_128B_BASE32B atom (4 k rows of 128 B, 32-byte chunk index xor-ed with the k row), atoms 512 B apart
    // bit 6: the same physical atom read as a K-major operand (K == 32: one 128 B row of k per n, 8 n rows = 1 KB)
    const uint32_t off = (variant & 64) ? (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u + (uint32_t)((((k >> 3) & 3) ^ (n & 3)) * 32) + (uint32_t)(k & 7) * 4u
                       : (variant & 32) ? (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u + (uint32_t)((((n >> 3) & 3) ^ (k & 3)) * 32) + (uint32_t)(n & 7) * 4u
                       : (variant & 4) ? (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u + (uint32_t)(((n >> 2) ^ (k & 7)) * 16) + (uint32_t)(n & 3) * 4u
                                       : tc::b_tile_offset(n, k, N);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t a_hi = tbase, a_lo = tbase + (uint32_t)K, d_col = tbase + 2u * (uint32_t)K;
  if (warp < 4) {
    const int row = tid;
    const uint32_t lane_addr = ((uint32_t)(32 * warp)) << 16;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tc::tf32_split(A[row * K + k0 + i], hi[i], lo[i]);
      tc::tmem_st8(a_hi + lane_addr + (uint32_t)k0, hi);
      tc::tmem_st8(a_lo + lane_addr + (uint32_t)k0, lo);
    }
    tc::wait_st();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4 && lane == 0) {
    tc::fence_after_sync();
    const uint32_t lbo = (variant & 1) ? 128u : (uint32_t)N * 16u;
    const uint32_t sbo = (variant & 1) ? (uint32_t)N * 16u : 128u;
    // variant bit 1: M = 64 (experiment: where do the 64 rows of D land in tensor memory?)
    if (variant & 4) {      // MN-major B, SWIZZLE_128B: descriptor layout type 2, SBO = 1 KB between 8-k atoms; modes in bits 3..4
      const uint32_t idesc = tc::idesc_tf32(128, N) | ((variant & 64) ? 0u : (1u << 16));
      const int mode = (variant >> 3) & 3;
      uint32_t lbo_s = (mode & 1) ? 1024u : 128u, sbo_s = (mode & 1) ? 128u : 1024u;
      if (variant & 32) { lbo_s = (mode & 1) ? 512u : 1024u; sbo_s = (mode & 1) ? 1024u : 512u; }
      if (variant & 64) { lbo_s = (mode & 1) ? 1024u : 16u; sbo_s = (mode & 1) ? 16u : 1024u; }
      const uint32_t k_adv = (variant & 64) ? 32u : 1024u;
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 0) ? a_lo : a_hi;
        const uint32_t b = tc::smem_u32((pass == 1) ? b_lo : b_hi);
        for (int s = 0; s < K / 8; ++s) {
          const uint64_t bd = tc::smem_desc(b + (uint32_t)s * k_adv, lbo_s, sbo_s) | ((uint64_t)((mode & 2) ? 1 : 2) << 61);
          tc::mma_tf32_ts(d_col, a + 8u * (uint32_t)s, bd, idesc, pass > 0 || s > 0);
        }
      }
    } else
    tc::mma_3xtf32(d_col, a_hi, a_lo, tc::smem_u32(b_hi), tc::smem_u32(b_lo), N, K / 8, tc::idesc_tf32((variant & 2) ? 64 : 128, N), false, lbo, sbo);
    tc::mma_commit(&bar);
  }
  if (warp < 4) {
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    const uint32_t lane_addr = ((uint32_t)(32 * warp)) << 16;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float v[8];
      tc::tmem_ld8(d_col + lane_addr + (uint32_t)n0, v);
      tc::wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) D[tid * N + n0 + i] = v[i];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// TMA + swizzled-operand building block of the gradient kernel: T is [R][128] fp32 (one 512-byte row of 128 samples per
// column, the checkpoint layout).  For each of the 4 sample blocks a TMA box (32 samples x R rows, 128-byte swizzle)
// lands in shared memory, a lo tile is formed element-wise and D[128 x N] += T[0..127][:] . T[rB..rB+N][:]' is accumulated
// through SS-mode tcgen05.mma with SWIZZLE_128B K-major descriptors (3 passes).  `raw` (nullable) receives the shared-memory
// image of the first block (R * 32 floats) so that the host can check the swizzle pattern.
__global__ void __launch_bounds__(128, 1) tma_selftest_kernel(const __grid_constant__ CUtensorMap tmap, int R, int rB, int N,
                                                              float* __restrict__ D, float* __restrict__ raw) {
  extern __shared__ float4 smem4[];
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  smem += (1024u - (tc::smem_u32(smem) & 1023u)) & 1023u;
  uint8_t* tH = smem;
  uint8_t* tL = smem + (uint32_t)R * 128u;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar_full, 1); tc::mbar_init(&bar_mma, 1); tc::mbar_fence_init(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  for (int sub = 0; sub < 4; ++sub) {
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(&bar_full, (uint32_t)R * 128u);
      tc::tma_load_3d(tH, &tmap, &bar_full, sub * 32, 0, 0);
    }
    tc::mbar_wait(&bar_full, (uint32_t)sub & 1u);
    if (sub == 0 && raw)
      for (int q = tid; q < R * 32; q += 128) raw[q] = reinterpret_cast<const float*>(tH)[q];
    for (int q = tid; q < R * 8; q += 128) reinterpret_cast<float4*>(tL)[q] = lo4(reinterpret_cast<const float4*>(tH)[q]);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t sH = tc::smem_u32(tH), sL = tc::smem_u32(tL), id = tc::idesc_tf32(128, N);
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 0) ? sL : sH, b = ((pass == 1) ? sL : sH) + (uint32_t)rB * 128u;
        for (int ks = 0; ks < 4; ++ks)
          tc::mma_tf32_ss(tbase, tc::smem_desc_sw128(a + (uint32_t)ks * 32u), tc::smem_desc_sw128(b + (uint32_t)ks * 32u), id,
                          sub > 0 || pass > 0 || ks > 0);
      }
      tc::mma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, (uint32_t)sub & 1u);       // the tile is overwritten by the next block
    tc::fence_after_sync();
  }
  const uint32_t lane_addr = ((uint32_t)(32 * warp)) << 16;
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    tc::tmem_ld8(tbase + lane_addr + (uint32_t)n0, v);
    tc::wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) D[tid * N + n0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace pspde
#endif

extern "C" int pspde_tma_selftest(int R, int rB, int N, const float* T, float* D, float* raw, void* stream) {
#if defined(PSPDE_EMULATE)
  (void)R; (void)rB; (void)N; (void)T; (void)D; (void)raw; (void)stream;
  return fail(-20, "the tensor-core path does not exist in the host emulator");
#else
  if (R < 128 || R > 256 || (R & 7) || (rB & 7) || rB < 0 || N < 16 || (N & 15) || N > 256 || rB + N > R || !T || !D)
    return fail(-2, "bad selftest shape");
  GradTcGeom tg;
  memset(&tg, 0, sizeof(tg));
  tg.cols = R; tg.box_rows = R;
  CUtensorMap tmap;
  if (grad_tc_tensor_map(tg, T, 1, &tmap)) return fail(-11, "cuTensorMapEncodeTiled failed");
  const size_t smem = 2 * (size_t)R * 128 + 1024;
  if (pspde_set_smem(tma_selftest_kernel, smem)) return fail(-11, "cudaFuncSetAttribute failed");
  tma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmap, R, rB, N, D, raw);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "tma_selftest launch failed: %s", e);
  return 0;
#endif
}

extern "C" int pspde_tc_selftest(int K, int N, int variant, const float* A, const float* B, float* D, void* stream) {
#if defined(PSPDE_EMULATE)
  (void)K; (void)N; (void)variant; (void)A; (void)B; (void)D; (void)stream;
  return fail(-20, "the tensor-core path does not exist in the host emulator");
#else
  if ((variant & 4) && (N != 32 || K > 32)) return fail(-2, "the swizzled-operand probe is N = 32, K <= 32");
  if (K < 8 || (K & 7) || N < 16 || (N & 15) || N > 256 || 2 * K + N > 512 || !A || !B || !D) return fail(-2, "bad selftest shape");
  const size_t smem = 2 * (size_t)tc::b_tile_bytes(K, N) + ((variant & 4) ? 4096 : 0);   // probe variants: 1 KB alignment + K-major atoms
  if (smem > kMaxSmem) return fail(-6, "selftest tile too large");
  if (pspde_set_smem(tc_selftest_kernel, smem)) return fail(-11, "cudaFuncSetAttribute failed");
  PSPDE_LAUNCH(tc_selftest_kernel, 1, 160, smem, stream, K, N, variant, A, B, D);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "tc_selftest launch failed: %s", e);
  return 0;
#endif
}

// ---- tcgen05 issue / completion cost probe (tools/probe_mma_cost.py): chains of n kind::tf32 MMAs on zeroed operands, timed
// with clock64 from the first issue to the arrival of the commit.  out[3 c + 0..2] = cycles, n, (M << 16 | N) for case c.
#if !defined(PSPDE_EMULATE)
namespace pspde {
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, unsigned long long* __restrict__ out) {
  extern __shared__ float4 smem4[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  smem += (1024u - (tc::smem_u32(smem) & 1023u)) & 1023u;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int q = tid; q < 96 * 1024 / 16; q += 128) reinterpret_cast<float4*>(smem)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, sb = tc::smem_u32(smem);
  if (tid == 0) {
    uint32_t ph = 0;
    int c = 0;
    // mode 0: SS K-major no-swizzle, one accumulator; 1: SS, two accumulators alternating; 2: TS (A in tensor memory), one
    // accumulator; 3: SS, four accumulators round-robin
    auto run = [&](int mode, int M, int N) {
      const uint32_t id = tc::idesc_tf32(M, N);
      const uint32_t lboA = (uint32_t)M * 16u, lboB = (uint32_t)N * 16u;
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        const uint32_t d = tb + (mode == 1 ? (uint32_t)(i & 1) * 256u : mode == 3 ? (uint32_t)(i & 3) * 128u : 0u);
        const uint64_t bd = tc::smem_desc(sb + 49152u + (uint32_t)(i % 6) * 2u * lboB, lboB, 128u);
        if (mode == 2) tc::mma_tf32_ts(d, tb + 256u + 8u * (uint32_t)(i & 7), bd, id, i > 0);
        else tc::mma_tf32_ss(d, tc::smem_desc(sb + (uint32_t)(i % 6) * 2u * lboA, lboA, 128u), bd, id, i > 0);
      }
      const long long t1 = clock64();
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, ph); ph ^= 1u;
      const long long t2 = clock64();
      out[4 * c + 0] = (unsigned long long)(t2 - t0); out[4 * c + 1] = (unsigned long long)(t1 - t0);
      out[4 * c + 2] = (unsigned long long)n; out[4 * c + 3] = ((unsigned long long)mode << 32) | ((unsigned long long)M << 16) | (unsigned long long)N;
      ++c;
    };
    // modes 4 (TS) / 5 (SS): the 8 descriptors of a k loop are built BEFORE the timed region, the issue loop is 8 MMAs unrolled
    auto run_pre = [&](int mode, int M, int N) {
      const uint32_t id = tc::idesc_tf32(M, N);
      const uint32_t lboA = (uint32_t)M * 16u, lboB = (uint32_t)N * 16u;
      uint64_t bd[8], ad[8];
      uint32_t aa[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bd[j] = tc::smem_desc(sb + 49152u + (uint32_t)(j % 6) * 2u * lboB, lboB, 128u);
        ad[j] = tc::smem_desc(sb + (uint32_t)(j % 6) * 2u * lboA, lboA, 128u);
        aa[j] = tb + 256u + 8u * (uint32_t)j;
      }
      const long long t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (mode == 4) tc::mma_tf32_ts(tb, aa[j], bd[j], id, (i + j) > 0);
          else tc::mma_tf32_ss(tb, ad[j], bd[j], id, (i + j) > 0);
        }
      }
      const long long t1 = clock64();
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, ph); ph ^= 1u;
      const long long t2 = clock64();
      out[4 * c + 0] = (unsigned long long)(t2 - t0); out[4 * c + 1] = (unsigned long long)(t1 - t0);
      out[4 * c + 2] = (unsigned long long)((n + 7) / 8 * 8); out[4 * c + 3] = ((unsigned long long)mode << 32) | ((unsigned long long)M << 16) | (unsigned long long)N;
      ++c;
    };
    for (int rep = 0; rep < 2; ++rep) {
      c = 0;
      run(0, 64, 32); run(2, 64, 32); run(2, 128, 176); run(0, 128, 176);
      run_pre(4, 64, 32); run_pre(4, 128, 64); run_pre(4, 128, 176); run_pre(4, 64, 136); run_pre(4, 128, 256);
      run_pre(5, 64, 32); run_pre(5, 64, 64); run_pre(5, 128, 64); run_pre(5, 128, 176); run_pre(5, 128, 256);
    }
    out[4 * c] = 0ull;
    out[4 * 31] = (unsigned long long)c;
  }
  __syncthreads();
  if (warp == 1) {     // modes 6 (TS) / 7 (SS): the WHOLE warp runs the issue loop, one elected lane issues (no divergence)
    int c = (int)out[4 * 31];
    uint32_t ph = 0;   // bar has completed an even number of phases per rep pair? recomputed below
    ph = (uint32_t)(2 * c) & 1u;
    auto run_w = [&](int mode, int M, int N) {
      const uint32_t id = tc::idesc_tf32(M, N);
      const uint32_t lboA = (uint32_t)M * 16u, lboB = (uint32_t)N * 16u;
      uint64_t bd[8], ad[8];
      uint32_t aa[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bd[j] = tc::smem_desc(sb + 49152u + (uint32_t)(j % 6) * 2u * lboB, lboB, 128u);
        ad[j] = tc::smem_desc(sb + (uint32_t)(j % 6) * 2u * lboA, lboA, 128u);
        aa[j] = tb + 256u + 8u * (uint32_t)j;
      }
      __syncwarp();
      const long long t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t acc = (i + j) > 0 ? 1u : 0u;
          if (mode == 6)
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                         "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tb), "r"(aa[j]), "l"(bd[j]), "r"(id), "r"(acc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                         "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tb), "l"(ad[j]), "l"(bd[j]), "r"(id), "r"(acc) : "memory");
        }
      }
      const long long t1 = clock64();
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                   "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(tc::smem_u32(&bar)) : "memory");
      tc::mbar_wait(&bar, ph); ph ^= 1u;
      const long long t2 = clock64();
      if ((tid & 31) == 0) {
        out[4 * c + 0] = (unsigned long long)(t2 - t0); out[4 * c + 1] = (unsigned long long)(t1 - t0);
        out[4 * c + 2] = (unsigned long long)((n + 7) / 8 * 8); out[4 * c + 3] = ((unsigned long long)mode << 32) | ((unsigned long long)M << 16) | (unsigned long long)N;
      }
      ++c;
    };
    run_w(6, 64, 32); run_w(6, 128, 176); run_w(6, 64, 136); run_w(7, 64, 32); run_w(7, 128, 176);
    if ((tid & 31) == 0) out[4 * c] = 0ull;
  }
  // modes 8 / 9: TWO issuing threads (warps 2 and 3, lane 0) at the same time, each n MMAs into its own accumulator
  // (8: both SS M = 64 N = 32; 9: warp 2 SS M = 64 N = 64, warp 3 TS M = 128 N = 176): is the ~50-cycle issue cost per thread?
  __shared__ uint64_t bar2[2];
  __shared__ int c_s;
  if (tid == 0) { tc::mbar_init(&bar2[0], 1); tc::mbar_init(&bar2[1], 1); tc::mbar_fence_init(); }
  __syncthreads();
  if (tid == 32) c_s = 0;
  for (int mode = 8; mode <= 9; ++mode) {
    __syncthreads();
    if ((warp == 2 || warp == 3) && (tid & 31) == 0) {
      const int w = warp - 2;
      const bool ts = mode == 9 && w == 1;
      const int M = ts ? 128 : 64, N = ts ? 176 : (mode == 9 ? 64 : 32);
      const uint32_t id = tc::idesc_tf32(M, N);
      const uint32_t lboA = (uint32_t)M * 16u, lboB = (uint32_t)N * 16u;
      uint64_t bd[8], ad[8];
      uint32_t aa[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bd[j] = tc::smem_desc(sb + 49152u + (uint32_t)(j % 6) * 2u * lboB, lboB, 128u);
        ad[j] = tc::smem_desc(sb + (uint32_t)(j % 6) * 2u * lboA, lboA, 128u);
        aa[j] = tb + 448u + 8u * (uint32_t)j;
      }
      const uint32_t d = tb + (uint32_t)w * 192u;
      const long long t0 = clock64();
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (ts) tc::mma_tf32_ts(d, aa[j], bd[j], id, (i + j) > 0);
          else tc::mma_tf32_ss(d, ad[j], bd[j], id, (i + j) > 0);
        }
      }
      const long long t1 = clock64();
      tc::mma_commit(&bar2[w]);
      tc::mbar_wait(&bar2[w], (uint32_t)(mode - 8) & 1u);
      const long long t2 = clock64();
      const int c = (int)out[4 * 31] + 5 + 2 * (mode - 8) + w;
      out[4 * c + 0] = (unsigned long long)(t2 - t0); out[4 * c + 1] = (unsigned long long)(t1 - t0);
      out[4 * c + 2] = (unsigned long long)((n + 7) / 8 * 8); out[4 * c + 3] = ((unsigned long long)mode << 32) | ((unsigned long long)M << 16) | (unsigned long long)N;
      if (mode == 9 && w == 1) out[4 * (c + 1)] = 0ull;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}
}  // namespace pspde
#endif

extern "C" int pspde_mma_probe(int n, unsigned long long* out, void* stream) {
#if defined(PSPDE_EMULATE)
  (void)n; (void)out; (void)stream;
  return fail(-20, "the tensor-core path does not exist in the host emulator");
#else
  if (n < 1 || n > 4096 || !out) return fail(-2, "bad probe arguments");
  const size_t smem = 97 * 1024 + 1024;
  if (pspde_set_smem(mma_probe_kernel, smem)) return fail(-11, "cudaFuncSetAttribute failed");
  mma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n, out);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "mma_probe launch failed: %s", e);
  return 0;
#endif
}
