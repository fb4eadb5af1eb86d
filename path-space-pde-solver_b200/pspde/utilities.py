"""Evaluation utilities on the fused path: ``do_importance_sampling_me`` with the reference's signature
(utilities.py:287-359).  Forward-only rollout of the CONTROLLED process u = -Z_n(X, t) on the evaluation grid
``delta_t`` (which may differ from the solver's), Girsanov weights, then mean / variance / relative error of
exp(-int f - g(X_T)) * exp(-ito - riemann / 2).  The rollout is one launch of the forward kernel
(``pspde_importance_sampling``); the three scalar statistics are formed in fp64 on the device."""
import ctypes

import numpy as np
import torch as pt

from . import _lib as L
from . import dist


def _time_index(model, N, delta_t):
    """n_net[n] = int(ceil(t / model.delta_t)) for t = n * delta_t, evaluated exactly like Solver.Z_n (solver.py:360-362):
    python float divided by the solver's 0-dim fp32 tensor."""
    dt_model = model.delta_t.detach().cpu()
    return pt.tensor([int(pt.ceil((n * delta_t) / dt_model)) for n in range(N)], dtype=pt.int32)


def do_importance_sampling_me(problem, model, K, control='approx', simulate_naive=False, verbose=False, delta_t=0.01,
                              on_cpu=False, cross_statistics=None, xis=None, seed=None):
    """Returns (mean_IS, variance_IS, rel_error_IS) like the reference.  `xis` (N, K, d) injects the increments
    (the reference draws pt.randn(K, d) per step, utilities.py:310); default: in-kernel Philox."""
    if control != 'approx' or simulate_naive or on_cpu:
        raise NotImplementedError("only control='approx' without the naive simulation is on the fused path")
    eng = model._get_engine()
    lib, dev = eng.lib, eng.device
    rank, W = dist.world(model.process_group)
    lo, hi = dist.shard_range(K, rank, W)
    Kl, d = hi - lo, problem.d
    N = int(np.ceil(problem.T / delta_t))
    t_index = _time_index(model, N, delta_t).to(dev)
    strides, noise, xi_ptr = (0, 0, 0), L.NOISE_PHILOX, None
    if xis is not None:
        xis = xis[:, lo:hi].to(dev, pt.float32).contiguous()
        strides, noise, xi_ptr = (d, 1, Kl * d), L.NOISE_INJECT, ctypes.c_void_p(xis.data_ptr())
    cfg = L.make_cfg(Kl, d, N, float(pt.tensor(delta_t, dtype=pt.float32)), eng.problem_id, eng.net_id, eng.dims,
                     eng.time_mode, adaptive=True, k_offset=lo, problem_flags=eng.flags, noise_mode=noise,
                     seed=(model.seed if seed is None else seed) + 0x15, offset=len(model.IS_rel_log),
                     xi_strides=strides, n_sets=(model.N if eng.time_mode == L.TIME_NONE else 0))
    f32 = dict(dtype=pt.float32, device=dev)
    Y, gX, F, X = pt.empty(Kl, **f32), pt.empty(Kl, **f32), pt.empty(Kl, **f32), pt.empty(Kl, d, **f32)
    ws = pt.empty(max(int(lib.pspde_workspace_bytes_fwd(ctypes.byref(cfg))), 4096), dtype=pt.uint8, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    # every path starts from problem.X_0 (utilities.py:302), also when training uses random_X_0 (eng.x0 is then per path)
    x0 = problem.X_0.detach().to(dev, pt.float32).contiguous()
    rc = lib.pspde_importance_sampling(ctypes.byref(cfg), p(model._theta.detach()), p(eng.pack), p(x0), xi_ptr,
                                       p(t_index), ctypes.c_float(float(model.delta_t)), p(X), p(Y), p(gX), p(F), p(ws),
                                       ws.numel(), ctypes.c_void_p(pt.cuda.current_stream(dev).cuda_stream))
    L.check(lib, rc)
    w = pt.exp(Y.double() - 2.0 * F.double() - gX.double())      # exp(-Fint - g) * exp(-ito - riemann / 2)
    ok = pt.isfinite(w)
    w = pt.where(ok, w, pt.zeros_like(w))
    s = pt.stack([w.sum(), (w * w).sum(), ok.sum().double()])
    dist.all_reduce_sum_(s, model.process_group)
    s1, s2, n = s.tolist()
    mean_IS = s1 / n
    variance_IS = (s2 - n * mean_IS * mean_IS) / (n - 1.0)        # pt.var: unbiased
    rel_error_IS = float(np.sqrt(max(variance_IS, 0.0)) / mean_IS)
    if verbose:
        string = 'IS mean: %.4e, IS variance: %.4e, IS RE %.4e' % (mean_IS, variance_IS, rel_error_IS)
        if cross_statistics is not None:
            crossed = dist.all_reduce_sum_((X > cross_statistics).sum().double().reshape(1), model.process_group)[0]
            string += ', crossed: %d/%d' % (int(crossed), K)
        print(string)
    return mean_IS, variance_IS, rel_error_IS
