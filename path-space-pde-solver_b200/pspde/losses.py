"""Path-space losses of Solver.loss_function (solver.py:164-192) as (value, dL/dY_N, dL/dZsum, dL/dg(X_N)) on the device.

All batch statistics are GLOBAL over the K_global trajectories of all ranks (one all_reduce of a few fp64 sums);
each rank gets the cotangents of its own paths.  Everything stays on the device: no host synchronisation.
"""
import torch as pt

from .dist import all_reduce_sum_

SUPPORTED = ("log-variance", "moment", "variance", "cross_entropy", "relative_entropy")


def value_and_cotangents(method, Y, gX, Zsum, K_global, adaptive=True, group=None, stats=None):
    """Y, gX, Zsum: (K_local,) fp32.  stats: optional fp64 [sum D, sum D^2, sum(Zsum+gX), #nonfinite] of the local
    shard as produced by the forward kernel (non-finite trajectories already excluded from the sums).
    Returns (loss fp64 0-dim, wY, wZ, wG fp32 or None, n_bad fp64 0-dim); wG = dL/dg(X_N) only matters when the
    forward process is attached (X_N depends on theta).

    Trajectories whose D = Y_N - g(X_N) is not finite (a rare blow-up of the untrained feedback control at large
    K) are dropped from the batch: zero cotangent, statistics over the K_eff = K - n_bad remaining ones.  The
    reference would return NaN for the whole batch in that case (NaNs propagate silently);
    with no such trajectory the formulas below are exactly solver.py:164-192."""
    D = (Y - gX).double()
    ok = pt.isfinite(D) & pt.isfinite(Zsum)
    nb = (~ok).sum().double().reshape(1)
    D = pt.where(ok, D, pt.zeros_like(D))
    zero = pt.zeros_like(D)

    def reduce(*sums):
        """ONE all-reduce per iteration: the method's fp64 sums and the count of dropped trajectories together."""
        s = all_reduce_sum_(pt.cat([pt.stack(list(sums)) if len(sums) > 1 else sums[0].reshape(-1), nb]), group)
        return s[:-1], s[-1], float(K_global) - s[-1]

    if method in ("log-variance", "moment"):
        s, n_bad, K = reduce(stats[:2].clone() if stats is not None else pt.stack([D.sum(), (D * D).sum()]))
        mean = s[0] / K
        if method == "moment":                                       # :165-166
            w = pt.where(ok, D * (2.0 / K), zero).float()
            return s[1] / K, w, None, -w, n_bad
        w = pt.where(ok, (D - mean) * (2.0 / K), zero).float()
        return s[1] / K - mean * mean, w, None, -w, n_bad               # :167-168 (biased variance)
    if method == "variance":                                         # :171-172  pt.var (unbiased) of exp(-g + Y)
        E = pt.where(ok, pt.exp(D), zero)
        s, n_bad, K = reduce(E.sum(), (E * E).sum())
        mean = s[0] / K
        w = pt.where(ok, 2.0 * (E - mean) * E / (K - 1.0), zero).float()
        return (s[1] - K * mean * mean) / (K - 1.0), w, None, -w, n_bad
    if method == "cross_entropy":                                    # :183-186
        E = pt.where(ok, pt.exp(D) if adaptive else pt.exp(-gX.double()), zero)
        s, n_bad, K = reduce((pt.where(ok, Y.double(), zero) * E).sum())
        return s[0] / K, (E / K).float(), None, (-pt.where(ok, Y.double(), zero) * E / K).float(), n_bad
    if method == "relative_entropy":                                 # :179-180 with a detached forward process
        s, n_bad, K = reduce(stats[2:3].clone() if stats is not None
                             else pt.where(ok, Zsum.double() + gX.double(), zero).sum())
        w = pt.where(ok, pt.ones_like(D) / K, zero).float()
        return s[0] / K, None, w, w, n_bad
    raise NotImplementedError("loss_method %r is not implemented by the fused solver (supported: %s)"
                              % (method, ", ".join(SUPPORTED)))
