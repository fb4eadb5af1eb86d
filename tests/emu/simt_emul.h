// simt_emul.h -- TEST-ONLY host emulation of the CUDA constructs the pspde kernels use.
//
// Compiles path-space-pde-solver_b200/csrc/*.cuh|*.cu for the CPU (-DPSPDE_EMULATE) so that index logic, smem
// layout and barrier placement can be checked against the oracle without a GPU.  Every CUDA thread is a
// ucontext fiber; __syncthreads / warp shuffles yield to a scheduler that releases a barrier once every live
// thread of the block (warp) has arrived.  Blocks run one after the other.  This file is never part of the
// product library (libpspde.so is built by nvcc only and fails loudly without a GPU).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_emu { unsigned x, y, z; };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu {

enum Wait { RUN = 0, BLOCK_BAR = 1, WARP_BAR = 2, DONE = 3 };

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  int tid = 0;
  int state = RUN;
};

struct Machine {
  ucontext_t main_ctx;
  std::vector<Fiber> fibers;
  Fiber* cur = nullptr;
  unsigned block_idx = 0, block_dim = 1, grid_dim = 1;
  std::vector<char> smem;
  char* smem_aligned = nullptr;
  uint64_t xchg[64][32];  // per-warp shuffle exchange buffer
  std::function<void()> body;
  int sm_count = 4;
  unsigned long long launches = 0;
};

inline Machine& M() { static Machine m; return m; }

inline void yield(int why) {
  Machine& m = M();
  m.cur->state = why;
  swapcontext(&m.cur->ctx, &m.main_ctx);
}

inline void trampoline() {
  Machine& m = M();
  m.body();
  m.cur->state = DONE;
  swapcontext(&m.cur->ctx, &m.main_ctx);
}

inline void run_block(unsigned bidx, unsigned bdim, unsigned gdim, size_t smem_bytes) {
  Machine& m = M();
  m.block_idx = bidx; m.block_dim = bdim; m.grid_dim = gdim;
  m.smem.assign(smem_bytes + 64, (char)0xCD);  // poison: uninitialised shared memory reads become visible
  m.smem_aligned = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(m.smem.data()) + 15) & ~uintptr_t(15));
  m.fibers.resize(bdim);
  for (unsigned t = 0; t < bdim; ++t) {
    Fiber& f = m.fibers[t];
    f.tid = (int)t; f.state = RUN;
    if (f.stack.size() < (1u << 18)) f.stack.resize(1u << 18);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = &m.main_ctx;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  // PSPDE_EMU_ORDER = forward (default) | reverse | random: the order in which runnable threads are resumed
  // between barriers.  A correct kernel gives identical results under every order; a missing barrier does not.
  static const char* order_env = getenv("PSPDE_EMU_ORDER");
  const int order = !order_env ? 0 : (order_env[0] == 'r' && order_env[1] == 'e') ? 1 : (order_env[0] == 'r') ? 2 : 0;
  std::vector<unsigned> perm(bdim);
  for (unsigned t = 0; t < bdim; ++t) perm[t] = t;
  uint64_t rng = 0x9E3779B97F4A7C15ull * (bidx + 1);
  for (;;) {
    bool any = false;
    if (order == 2) {  // shuffle at WARP granularity plus lanes inside a warp
      for (unsigned t = bdim - 1; t > 0; --t) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        unsigned u = (unsigned)(rng % (t + 1));
        unsigned tmp = perm[t]; perm[t] = perm[u]; perm[u] = tmp;
      }
    }
    for (unsigned i = 0; i < bdim; ++i) {
      const unsigned t = order == 1 ? bdim - 1 - i : perm[i];
      Fiber& f = m.fibers[t];
      if (f.state == RUN) { any = true; m.cur = &f; swapcontext(&m.main_ctx, &f.ctx); }
    }
    // release barriers
    bool all_done = true, all_block = true;
    for (auto& f : m.fibers) { if (f.state != DONE) { all_done = false; if (f.state != BLOCK_BAR) all_block = false; } }
    if (all_done) break;
    bool released = false;
    if (all_block) { for (auto& f : m.fibers) if (f.state == BLOCK_BAR) f.state = RUN; released = true; }
    else {
      for (unsigned w = 0; w * 32 < bdim; ++w) {
        bool ok = true, anyw = false;
        for (unsigned l = 0; l < 32 && w * 32 + l < bdim; ++l) {
          int s = m.fibers[w * 32 + l].state;
          if (s == WARP_BAR) anyw = true; else if (s != DONE) ok = false;
        }
        if (ok && anyw) { for (unsigned l = 0; l < 32 && w * 32 + l < bdim; ++l) if (m.fibers[w * 32 + l].state == WARP_BAR) m.fibers[w * 32 + l].state = RUN; released = true; }
      }
    }
    (void)any;
    if (!released) { fprintf(stderr, "emu: deadlock (divergent barrier) in block %u\n", bidx); abort(); }
  }
}

template <typename F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& f) {
  Machine& m = M();
  m.body = f;
  m.launches++;
  for (unsigned b = 0; b < grid.x; ++b) run_block(b, block.x, grid.x, smem_bytes);
}

inline uint3_emu thread_idx() { return uint3_emu{(unsigned)M().cur->tid, 0, 0}; }
inline uint3_emu block_idx() { return uint3_emu{M().block_idx, 0, 0}; }
inline uint3_emu block_dim() { return uint3_emu{M().block_dim, 1, 1}; }
inline uint3_emu grid_dim() { return uint3_emu{M().grid_dim, 1, 1}; }

template <typename T>
inline T shfl_xor(T v, int lanemask) {
  static_assert(sizeof(T) <= 8, "shuffle type");
  Machine& m = M();
  const int tid = m.cur->tid, w = tid >> 5, l = tid & 31;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  m.xchg[w][l] = bits;
  yield(WARP_BAR);
  uint64_t got = m.xchg[w][(l ^ lanemask) & 31];
  yield(WARP_BAR);
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::block_idx())
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())
#define PSPDE_DYN_SMEM(name) float4* name = reinterpret_cast<float4*>(emu::M().smem_aligned)

inline void __syncthreads() { emu::yield(emu::BLOCK_BAR); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield(emu::WARP_BAR); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lanemask) { return emu::shfl_xor(v, lanemask); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __logf(float x) { return logf(x); }
inline void __sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
using std::isfinite;
