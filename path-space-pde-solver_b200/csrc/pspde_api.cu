// pspde_api.cu -- extern "C" entry points declared in include/pspde.h: validation, launch planning, launches.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/pspde.h"
#include "rollout_kernels.cuh"

#if defined(PSPDE_EMULATE)
#define PSPDE_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
static inline int pspde_sm_count() { return emu::M().sm_count; }
static inline int pspde_memset0(void* p, size_t n, void*) { memset(p, 0, n); return 0; }
template <typename K> static inline int pspde_set_smem(K, size_t) { return 0; }
static inline const char* pspde_peek_error() { return nullptr; }
#else
#define PSPDE_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
static inline int pspde_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { n = 0; return -1; }
  }
  return n;
}
static inline int pspde_memset0(void* p, size_t n, void* stream) {
  return cudaMemsetAsync(p, 0, n, (cudaStream_t)stream) == cudaSuccess ? 0 : -1;
}
template <typename K> static inline int pspde_set_smem(K kern, size_t bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -1;
}
static inline const char* pspde_peek_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
#endif

using namespace pspde;

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

constexpr int kP = 64;                       // trajectories per tile
constexpr size_t kMaxSmem = 227 * 1024;      // opt-in dynamic shared memory per CTA on sm_100

struct Plan {
  NetGeom g;
  int T, NB, n_tiles, grid, n_sets, n_theta_total;
  int r_fwd[PSPDE_MAXL];
  size_t smem_bytes, stats_bytes, grad_bytes;
};

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int validate(const pspde_cfg* c) {
  if (!c) return fail(-1, "cfg is NULL");
  if (c->K_local < 1 || c->d < 1 || c->N < 0) return fail(-2, "bad sizes K_local=%d d=%d N=%d", c->K_local, c->d, c->N);
  if (!(c->dt > 0.f)) return fail(-2, "dt must be > 0");
  if (c->n_layers < 1 || c->n_layers > PSPDE_MAX_LAYERS) return fail(-3, "n_layers=%d unsupported (1..%d)", c->n_layers, PSPDE_MAX_LAYERS);
  if (c->net_id != PSPDE_NET_DENSENET && c->net_id != PSPDE_NET_MLP_TANH) return fail(-3, "unknown net_id %d", c->net_id);
  if (c->time_mode < 0 || c->time_mode > 2) return fail(-3, "unknown time_mode %d", c->time_mode);
  if (c->problem_id != PSPDE_PROBLEM_OU && c->problem_id != PSPDE_PROBLEM_DW)
    return fail(-4, "problem_id %d is not supported by the HJB rollout", c->problem_id);
  if ((c->problem_flags & PSPDE_FLAG_DENSE_AB) && c->d > 128) return fail(-4, "dense A/B needs d <= 128 (d=%d)", c->d);
  if (c->dims[c->n_layers] != c->d) return fail(-3, "control network must map to d outputs (got %d)", c->dims[c->n_layers]);
  if (c->noise_mode != PSPDE_NOISE_INJECT && c->noise_mode != PSPDE_NOISE_PHILOX) return fail(-5, "unknown noise_mode");
  return 0;
}

// paths per thread tile in gemm_nn: minimise (waves over the CTA) x (issue slots per 4-wide k step)
int choose_r(int Np, int T, int rmax) {
  int best = 1;
  long best_cost = -1;
  for (int R = 1; R <= rmax; R *= 2) {
    const long tiles = (long)(kP / R) * (Np / 4);
    const long waves = (tiles + T - 1) / T;
    const long cost = waves * (R + 4 + 16 * R);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = R; }
  }
  return best;
}

int make_plan(const pspde_cfg* c, bool bwd, bool attached, Plan& pl) {
  int rc = validate(c);
  if (rc) return rc;
  rc = build_geom(pl.g, c->net_id, c->n_layers, c->dims, c->time_mode, c->d);
  if (rc) return fail(-3, "network geometry rejected (code %d): dims[0] must be d%s", rc, c->time_mode == PSPDE_TIME_NONE ? "" : "+1");
  pl.n_sets = (c->time_mode == PSPDE_TIME_NONE) ? c->N : 1;
  pl.n_theta_total = pl.g.n_params * pl.n_sets;
  pl.n_tiles = (c->K_local + kP - 1) / kP;
  const int sms = pspde_sm_count();
  if (sms <= 0) return fail(-10, "no CUDA device");
  pl.grid = pl.n_tiles < sms ? pl.n_tiles : sms;
  pl.T = 512; pl.NB = 1;
  if (bwd) {
    const int nb = pl.g.n_blocks;
    if (nb <= 256) { pl.T = 256; pl.NB = 1; }
    else if (nb <= 512) { pl.T = 256; pl.NB = 2; }
    else if (nb <= 768) { pl.T = 256; pl.NB = 3; }
    else if (nb <= 1024) { pl.T = 512; pl.NB = 2; }
    else if (nb <= 1536) { pl.T = 512; pl.NB = 3; }
    else if (nb <= 2048) { pl.T = 256; pl.NB = 8; }
    else return fail(-6, "network too large for the register-resident gradient path (%d blocks > 2048)", nb);
  }
  for (int l = 0; l < pl.g.L; ++l) pl.r_fwd[l] = choose_r(pl.g.layer[l].Np, pl.T, bwd ? 4 : 8);
  const SmemLayout sl = smem_layout(pl.g, kP, bwd, attached);
  pl.smem_bytes = (size_t)sl.total * sizeof(float);
  if (pl.smem_bytes > kMaxSmem)
    return fail(-6, "network + tile need %zu B of shared memory (> %zu)", pl.smem_bytes, kMaxSmem);
  pl.stats_bytes = align256((size_t)pl.grid * 4 * sizeof(double));
  pl.grad_bytes = bwd ? align256((size_t)pl.grid * pl.n_theta_total * sizeof(float)) : 0;
  return 0;
}

void fill_params(const pspde_cfg* c, const Plan& pl, RolloutParams& p) {
  memset(&p, 0, sizeof(p));
  p.g = pl.g;
  p.K_local = c->K_local; p.k_offset = c->k_offset; p.d = c->d; p.N = c->N; p.dt = c->dt;
  p.problem_id = c->problem_id; p.flags = c->problem_flags; p.adaptive = c->adaptive;
  p.noise_mode = c->noise_mode; p.x0_per_path = c->x0_per_path;
  p.seed = c->seed; p.offset = c->offset;
  p.xs_k = c->xi_stride_k; p.xs_j = c->xi_stride_j; p.xs_n = c->xi_stride_n;
  p.n_tiles = pl.n_tiles; p.n_theta_total = pl.n_theta_total;
  for (int l = 0; l < PSPDE_MAXL; ++l) p.r_fwd[l] = pl.r_fwd[l];
}

template <int T, bool BWD, int NB>
int launch_rollout(const Plan& pl, const RolloutParams& p, void* stream) {
  auto kern = rollout_kernel<kP, T, BWD, NB>;
  if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute(%zu B smem) failed", pl.smem_bytes);
  PSPDE_LAUNCH(kern, pl.grid, T, pl.smem_bytes, stream, p);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "rollout kernel launch failed: %s", e);
  return 0;
}

}  // namespace

extern "C" {

int pspde_abi_version(void) { return PSPDE_ABI_VERSION; }
const char* pspde_last_error(void) { return g_err; }
uint64_t pspde_launch_count(void) { return g_launches.load(); }

int64_t pspde_theta_size(const pspde_cfg* cfg) {
  if (validate(cfg)) return -1;
  NetGeom g;
  if (build_geom(g, cfg->net_id, cfg->n_layers, cfg->dims, cfg->time_mode, cfg->d)) { fail(-3, "bad network geometry"); return -1; }
  return (int64_t)g.n_params * (cfg->time_mode == PSPDE_TIME_NONE ? cfg->N : 1);
}

size_t pspde_workspace_bytes(const pspde_cfg* cfg) {
  // upper bound over the entry points that are feasible for cfg (0 if none is)
  Plan pl;
  size_t need = 0;
  if (make_plan(cfg, false, false, pl) == 0) need = pl.stats_bytes;
  if (make_plan(cfg, true, false, pl) == 0) need = pl.stats_bytes + pl.grad_bytes;
  if (make_plan(cfg, true, true, pl) == 0)
    need = pl.stats_bytes + pl.grad_bytes + align256((size_t)pl.grid * cfg->N * kP * cfg->d * sizeof(float));
  return need ? need + 256 : 0;
}

int pspde_rollout_fwd(const pspde_cfg* cfg, const float* theta, const float* prob, const float* x0, const float* y0,
                      const float* xi, float* X_N, float* Y_N, float* gX, float* Zsum, double* stats,
                      void* workspace, size_t workspace_bytes, void* stream) {
  Plan pl;
  int rc = make_plan(cfg, false, false, pl);
  if (rc) return rc;
  if (!theta || !prob || !x0) return fail(-1, "theta/prob/x0 must not be NULL");
  if (cfg->noise_mode == PSPDE_NOISE_INJECT && !xi) return fail(-1, "noise_mode INJECT needs xi");
  if (!workspace || workspace_bytes < pl.stats_bytes) return fail(-7, "workspace too small (%zu < %zu)", workspace_bytes, pl.stats_bytes);
  RolloutParams p;
  fill_params(cfg, pl, p);
  p.theta = theta; p.prob = prob; p.x0 = x0; p.y0 = y0; p.xi = xi;
  p.X_N = X_N; p.Y_N = Y_N; p.gX = gX; p.Zsum = Zsum;
  p.stats_partial = reinterpret_cast<double*>(workspace);
  rc = launch_rollout<512, false, 1>(pl, p, stream);
  if (rc) return rc;
  if (stats) {
    PSPDE_LAUNCH(reduce_stats_kernel, 1, 32, 0, stream, p.stats_partial, pl.grid, stats);
    g_launches++;
    if (const char* e = pspde_peek_error()) return fail(-12, "reduce_stats launch failed: %s", e);
  }
  return 0;
}

int pspde_rollout_bwd_detached(const pspde_cfg* cfg, const float* theta, const float* prob, const float* x0,
                               const float* xi, const float* wY, const float* wZ, float* grad_theta,
                               void* workspace, size_t workspace_bytes, void* stream) {
  Plan pl;
  int rc = make_plan(cfg, true, false, pl);
  if (rc) return rc;
  if (!theta || !prob || !x0 || !wY || !grad_theta) return fail(-1, "theta/prob/x0/wY/grad_theta must not be NULL");
  if (cfg->noise_mode == PSPDE_NOISE_INJECT && !xi) return fail(-1, "noise_mode INJECT needs xi");
  if (!workspace || workspace_bytes < pl.stats_bytes + pl.grad_bytes)
    return fail(-7, "workspace too small (%zu < %zu)", workspace_bytes, pl.stats_bytes + pl.grad_bytes);
  RolloutParams p;
  fill_params(cfg, pl, p);
  p.theta = theta; p.prob = prob; p.x0 = x0; p.xi = xi; p.wY = wY; p.wZ = wZ;
  p.grad_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + pl.stats_bytes);
  if (pspde_memset0(p.grad_partial, (size_t)pl.grid * pl.n_theta_total * sizeof(float), stream))
    return fail(-12, "memset of the gradient partials failed");
  if (pl.T == 256 && pl.NB == 1) rc = launch_rollout<256, true, 1>(pl, p, stream);
  else if (pl.T == 256 && pl.NB == 2) rc = launch_rollout<256, true, 2>(pl, p, stream);
  else if (pl.T == 256 && pl.NB == 3) rc = launch_rollout<256, true, 3>(pl, p, stream);
  else if (pl.T == 256 && pl.NB == 8) rc = launch_rollout<256, true, 8>(pl, p, stream);
  else if (pl.T == 512 && pl.NB == 2) rc = launch_rollout<512, true, 2>(pl, p, stream);
  else if (pl.T == 512 && pl.NB == 3) rc = launch_rollout<512, true, 3>(pl, p, stream);
  else rc = fail(-13, "internal: no kernel for T=%d NB=%d", pl.T, pl.NB);
  if (rc) return rc;
  const int n = pl.n_theta_total;
  PSPDE_LAUNCH(reduce_grad_kernel, (n + 255) / 256, 256, 0, stream, p.grad_partial, pl.grid, n, grad_theta);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "reduce_grad launch failed: %s", e);
  return 0;
}

int pspde_rollout_attached(const pspde_cfg* cfg, const float* theta, const float* prob, const float* x0,
                           const float* xi, float w, float* X_N, float* gX, float* Zsum, double* stats,
                           float* grad_theta, void* workspace, size_t workspace_bytes, void* stream) {
  Plan pl;
  int rc = make_plan(cfg, true, true, pl);
  if (rc) return rc;
  if (!cfg->adaptive) return fail(-4, "attached mode implies adaptive_forward_process (solver.py:61-62)");
  if (!theta || !prob || !x0 || !grad_theta) return fail(-1, "theta/prob/x0/grad_theta must not be NULL");
  if (cfg->noise_mode == PSPDE_NOISE_INJECT && !xi) return fail(-1, "noise_mode INJECT needs xi");
  const size_t ckpt_bytes = align256((size_t)pl.grid * cfg->N * kP * cfg->d * sizeof(float));
  if (!workspace || workspace_bytes < pl.stats_bytes + pl.grad_bytes + ckpt_bytes)
    return fail(-7, "workspace too small (%zu < %zu)", workspace_bytes, pl.stats_bytes + pl.grad_bytes + ckpt_bytes);
  RolloutParams p;
  fill_params(cfg, pl, p);
  p.theta = theta; p.prob = prob; p.x0 = x0; p.xi = xi; p.w_attached = w;
  p.X_N = X_N; p.gX = gX; p.Zsum = Zsum;
  char* ws = reinterpret_cast<char*>(workspace);
  p.stats_partial = reinterpret_cast<double*>(ws);
  p.grad_partial = reinterpret_cast<float*>(ws + pl.stats_bytes);
  p.x_ckpt = reinterpret_cast<float*>(ws + pl.stats_bytes + pl.grad_bytes);
  if (pspde_memset0(p.grad_partial, (size_t)pl.grid * pl.n_theta_total * sizeof(float), stream))
    return fail(-12, "memset of the gradient partials failed");
#define PSPDE_ATT(TT, NBB)                                                                          \
  {                                                                                                 \
    auto kern = rollout_attached_kernel<kP, TT, NBB>;                                               \
    if (pspde_set_smem(kern, pl.smem_bytes)) return fail(-11, "cudaFuncSetAttribute failed");       \
    PSPDE_LAUNCH(kern, pl.grid, TT, pl.smem_bytes, stream, p);                                      \
  }
  if (pl.T == 256 && pl.NB == 1) PSPDE_ATT(256, 1)
  else if (pl.T == 256 && pl.NB == 2) PSPDE_ATT(256, 2)
  else if (pl.T == 256 && pl.NB == 3) PSPDE_ATT(256, 3)
  else if (pl.T == 256 && pl.NB == 8) PSPDE_ATT(256, 8)
  else if (pl.T == 512 && pl.NB == 2) PSPDE_ATT(512, 2)
  else if (pl.T == 512 && pl.NB == 3) PSPDE_ATT(512, 3)
  else return fail(-13, "internal: no kernel for T=%d NB=%d", pl.T, pl.NB);
#undef PSPDE_ATT
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "attached kernel launch failed: %s", e);
  if (stats) {
    PSPDE_LAUNCH(reduce_stats_kernel, 1, 32, 0, stream, p.stats_partial, pl.grid, stats);
    g_launches++;
  }
  const int n = pl.n_theta_total;
  PSPDE_LAUNCH(reduce_grad_kernel, (n + 255) / 256, 256, 0, stream, p.grad_partial, pl.grid, n, grad_theta);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "reduce launch failed: %s", e);
  return 0;
}

int pspde_philox_dump(const pspde_cfg* cfg, float* xi_out, void* stream) {
  if (!cfg || !xi_out) return fail(-1, "NULL argument");
  if (cfg->K_local < 1 || cfg->d < 1 || cfg->N < 1) return fail(-2, "bad sizes");
  PSPDE_LAUNCH(philox_dump_kernel, 64, 256, 0, stream, cfg->K_local, cfg->k_offset, cfg->d, cfg->N, cfg->seed,
               cfg->offset, xi_out);
  g_launches++;
  if (const char* e = pspde_peek_error()) return fail(-12, "philox_dump launch failed: %s", e);
  return 0;
}

int64_t pspde_fma_probe(int iters, float* sink, void* stream) {
  const int sms = pspde_sm_count();
  if (sms <= 0 || iters < 1 || !sink) { fail(-1, "bad arguments"); return -1; }
  PSPDE_LAUNCH(fma_probe_kernel, sms, 1024, 0, stream, iters, sink);
  g_launches++;
  if (const char* e = pspde_peek_error()) { fail(-12, "fma_probe launch failed: %s", e); return -1; }
  return (int64_t)sms * 1024 * (int64_t)iters * 16 * 8 * 2;
}

}  // extern "C"
