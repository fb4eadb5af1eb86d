// tc_sm100.cuh -- thin inline-PTX layer over the sm_100a tensor-core path (tcgen05 + tensor memory).
//
// What the rollout kernels need from the 5th-generation tensor cores is an FP32-equivalent product of a
// [128 paths x K] activation tile with a [K x N] weight tile.  kind::tf32 keeps 10 mantissa bits per operand, so
// each FP32 operand is split x = hi + lo with hi = cvt.rna.tf32(x) and lo = x - hi (exact in FP32) and the product
// is accumulated in three passes, hi*hi + lo*hi + hi*lo, in the FP32 accumulator in tensor memory (the dropped
// lo*lo term and the truncation of lo are O(2^-22) relative).
//   A operand: tensor memory, lane = row (path), one 32-bit column per k (written by the thread that owns the row
//              with tcgen05.st, so the activations never take a shared-memory layout).
//   B operand: shared memory, K-major, no swizzle: element (n, k) at  (k >> 2) * LBO + n * 16 + (k & 3) * 4  bytes,
//              i.e. one float4 per (k4, n) and 8 consecutive n form one 128-byte core matrix (SBO = 128 B,
//              LBO = 16 * N_pad bytes).  One tcgen05.mma consumes K = 8 (two core matrices along K).
//   D:         tensor memory, lane = row, column = n, FP32.
// Never compiled for the host emulator (tests/emu): the planning code falls back to the FP32-FMA kernels there.
#pragma once
#if !defined(PSPDE_EMULATE)
#include <cuda_runtime.h>
#include <stdint.h>

namespace pspde {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LAB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE_%=;\n\t"
      "bra LAB_WAIT_%=;\n\t"
      "LAB_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// non-blocking probe + blocking wait with a suspend-time hint (consumer side of a TMA / MMA pipeline)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// one arrival that also announces `bytes` of asynchronous (TMA) traffic that will complete_tx on this barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- TMA (cp.async.bulk.tensor): one box of a tiled tensor map -> shared memory, completion on an mbarrier.
// `tmap` is the generic address of a CUtensorMap that lives in kernel parameter space (__grid_constant__).
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// plain bulk copy global -> shared (no tensor map): `bytes` % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- tensor memory management (one warp allocates and later frees; the address lands in shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads B through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at [16,30),
// SBO >> 4 at [32,46), version = 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// K-major operand in the 128-byte-swizzle layout a TMA load with CU_TENSOR_MAP_SWIZZLE_128B produces: one row (M or N
// index) = 128 contiguous bytes = 32 tf32 values of K, rows 128 B apart, 8 rows = one 1 KB atom (16-byte chunk index
// xor-ed with the row index inside the atom), atoms of consecutive 8-row groups SBO = 1 KB apart; the tile must be 1 KB
// aligned.  A K = 8 MMA reads 32 B of every row: advance the start address by 32 B per k step inside the 128 B row.
// Layout type 2 (SWIZZLE_128B); LBO is not used by swizzled K-major operands (encoded as 1).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor for kind::tf32, FP32 accumulate, K-major A and B (cute::UMMA::InstrDescriptor):
// c_format = 1 at [4,6), a_format = b_format = 2 (TF32) at [7,10) / [10,13), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA: D[tmem] (+)= A[tmem] . B[smem]   (issued by ONE thread)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]   (issued by ONE thread)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all previously issued MMAs of this thread -> one arrival on `bar` when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- tensor memory <-> registers, 32 lanes x 32 bit: lane i of the warp <-> TMEM lane (taddr.lane + i), register j
// <-> column (taddr.col + j).  A warp may only touch the lane quarter 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
               "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
               "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
               "r"(__float_as_uint(v[15])) : "memory");
}

// ---- FP32 -> (hi, lo) TF32 split
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// hi = x truncated to TF32 -- exactly what the tensor core keeps of a raw FP32 operand -- and lo = x - hi (exact; x == hi + lo).
// (cvt.rna.tf32 would halve |lo|, but costs ~15 cycles of a worker's time per value where AND + FADD cost two; the dropped
// lo.lo term is 2^-21 instead of 2^-22 relative, the truncation of lo itself < 2^-21 |x|.)
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
// round-to-nearest split (|lo| <= 2^-11 |x|): for operands that are split once per kernel (the weights)
__device__ __forceinline__ void tf32_split_rn(float x, float& hi, float& lo) { hi = tf32_hi(x); lo = x - hi; }

// byte offset of element (n, k) inside a B tile of padded width Np (see the header comment)
__host__ __device__ constexpr uint32_t b_tile_offset(int n, int k, int Np) {
  return (uint32_t)(k >> 2) * (uint32_t)Np * 16u + (uint32_t)n * 16u + (uint32_t)(k & 3) * 4u;
}
__host__ __device__ constexpr uint32_t b_tile_bytes(int Kp, int Np) { return (uint32_t)(Kp >> 2) * (uint32_t)Np * 16u; }

// Issues the three passes of D[128 x N] (+)= A . B for K = 8 * ksteps (one thread).
//   a_hi / a_lo: TMEM addresses of column 0 of the hi / lo copies of A;  b_hi / b_lo: shared addresses of the tiles.
__device__ __forceinline__ void mma_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           int Np, int ksteps, uint32_t idesc, bool accumulate, uint32_t lbo, uint32_t sbo) {
  // small terms first: they are added while the accumulator is still small
  // one descriptor per pass; only its 14-bit start address (>> 4) moves from one k step to the next (a tcgen05.mma costs its
  // issuing thread 50 - 100 cycles, most of the excess being descriptor arithmetic in the loop: pspde_mma_probe)
  const uint64_t step = (uint64_t)((2u * (uint32_t)Np * 16u) >> 4);
  const uint64_t d_hi = smem_desc(b_hi, lbo, sbo), d_lo = smem_desc(b_lo, lbo, sbo);
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = (pass == 0) ? a_lo : a_hi;
    uint64_t bd = (pass == 1) ? d_lo : d_hi;
    for (int s = 0; s < ksteps; ++s, bd += step)
      mma_tf32_ts(d_tmem, a + 8u * (uint32_t)s, bd, idesc, accumulate || pass > 0 || s > 0);
  }
}

}  // namespace tc
}  // namespace pspde
#endif  // !PSPDE_EMULATE
