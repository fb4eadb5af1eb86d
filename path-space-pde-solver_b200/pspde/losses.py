"""Path-space losses of Solver.loss_function (solver.py:164-192) as (value, dL/dY_N, dL/dZsum) on the device.

All batch statistics are GLOBAL over the K_global trajectories of all ranks (one all_reduce of a few fp64 sums);
each rank gets the cotangents of its own paths.  Everything stays on the device: no host synchronisation.
"""
import torch as pt

from .dist import all_reduce_sum_

SUPPORTED = ("log-variance", "moment", "variance", "cross_entropy", "relative_entropy")


def value_and_cotangents(method, Y, gX, Zsum, K_global, adaptive=True, group=None, stats=None):
    """Y, gX, Zsum: (K_local,) fp32.  stats: optional fp64 [sum D, sum D^2, sum(Zsum+gX), #nonfinite] of the local
    shard as produced by the forward kernel.  Returns (loss fp64 0-dim, wY fp32 or None, wZ fp32 or None)."""
    K = float(K_global)
    D = (Y - gX).double()
    if method in ("log-variance", "moment"):
        s = stats[:2].clone() if stats is not None else pt.stack([D.sum(), (D * D).sum()])
        all_reduce_sum_(s, group)
        mean = s[0] / K
        if method == "moment":                                       # :165-166
            return s[1] / K, (D * (2.0 / K)).float(), None
        return s[1] / K - mean * mean, ((D - mean) * (2.0 / K)).float(), None   # :167-168 (biased variance)
    if method == "variance":                                         # :171-172  pt.var (unbiased) of exp(-g + Y)
        E = pt.exp(D)
        s = all_reduce_sum_(pt.stack([E.sum(), (E * E).sum()]), group)
        mean = s[0] / K
        return (s[1] - K * mean * mean) / (K - 1.0), (2.0 * (E - mean) * E / (K - 1.0)).float(), None
    if method == "cross_entropy":                                    # :183-186
        E = pt.exp(D) if adaptive else pt.exp(-gX.double())
        s = all_reduce_sum_((Y.double() * E).sum().reshape(1), group)
        return s[0] / K, (E / K).float(), None
    if method == "relative_entropy":                                 # :179-180 with a detached forward process
        s = stats[2:3].clone() if stats is not None else (Zsum.double() + gX.double()).sum().reshape(1)
        all_reduce_sum_(s, group)
        return s[0] / K, None, pt.full_like(Zsum, 1.0 / K)
    raise NotImplementedError("loss_method %r is not implemented by the fused solver (supported: %s)"
                              % (method, ", ".join(SUPPORTED)))
