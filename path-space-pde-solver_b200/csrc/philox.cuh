// philox.cuh -- Philox4x32-10 + Box-Muller, the in-kernel replacement for the reference's CPU-side
// xi = randn(K, d, N+1) (solver.py:381).  Counter = (k_global, n, j/4, offset), key = seed (include/pspde.h);
// oracle/philox.py is the bit-exact restatement of the integer part.
#pragma once
#include "simt.h"

namespace pspde {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1, unsigned (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// u = ((r >> 8) + 0.5) * 2^-24 lies strictly inside (0, 1): no log(0), |z| <= 5.9
__device__ __forceinline__ float u01(unsigned r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-8f; }

// sqrt.approx (one MUFU, ~1 ulp) -- the argument comes from __logf (2 ulp) anyway; sqrtf's Newton step costs ~10 instructions
__device__ __forceinline__ float sqrt_fast(float x) {
#if defined(PSPDE_EMULATE)
  return sqrtf(x);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

__device__ __forceinline__ float4 philox_normal4(unsigned k_global, unsigned n, unsigned jb, unsigned offset,
                                                 unsigned long long seed) {
  unsigned r[4];
  philox4x32_10(k_global, n, jb, offset, (unsigned)(seed & 0xffffffffull), (unsigned)(seed >> 32), r);
  float4 z;
  float s, c;
  const float rad0 = sqrt_fast(-2.0f * __logf(u01(r[0])));
  __sincosf(6.283185307179586f * u01(r[1]), &s, &c);
  z.x = rad0 * c; z.y = rad0 * s;
  const float rad1 = sqrt_fast(-2.0f * __logf(u01(r[2])));
  __sincosf(6.283185307179586f * u01(r[3]), &s, &c);
  z.z = rad1 * c; z.w = rad1 * s;
  return z;
}

}  // namespace pspde
