import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt, numpy as np
import pspde
from pspde.fused import Call

def mk(d, K, T=1.0, dt=0.01):
    prob = pspde.LLGC(d=d, off_diag=0, T=T, seed=42, device="cuda")
    S = pspde.Solver("c2", prob, K=K, L=1, delta_t=dt, time_approx="inner", detach_forward=True,
                     u_l2_error_flag=False, early_stopping_time=None, verbose=False)
    S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, seed=42)
    S.update_Phis()
    return S

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode in ("all", "fwd"):
    for d, K in ((10, 64 * 148), (10, 64 * 148 * 3), (100, 64 * 148), (100, 64 * 148 * 2), (100, 1 << 16)):
        S = mk(d, K); eng = S._get_engine(); theta = S._theta.detach()
        outs = []
        for rep in range(3):
            eng.forward(theta, None, Call(offset=0)); pt.cuda.synchronize()
            outs.append((eng.Y_N.clone(), eng.X_N.clone(), eng.stats.clone()))
        Y0 = outs[0][0]
        for rep in (1, 2):
            Y = outs[rep][0]
            neq = (Y != Y0) & ~(pt.isnan(Y) & pt.isnan(Y0))
            idx = neq.nonzero().flatten()
            print("d=%d K=%d rep%d: nonfinite=%d mismatches=%d maxdiff=%.3e first idx=%s tiles=%s stats3=%s" % (
                d, K, rep, int((~pt.isfinite(Y)).sum()), int(neq.sum()),
                float((Y - Y0)[neq].abs().max()) if neq.any() else 0.0, idx[:8].tolist(),
                sorted(set((idx // 64).tolist()))[:12], outs[rep][2].tolist()))
if mode in ("all", "train"):
    S = mk(100, 1 << 16); S.L = 8
    eng = S._get_engine()
    for l in range(8):
        S.train_step(l)
        g = S._theta.grad
        print("iter", l, "loss", S.loss_log[-1], "stats", eng.stats.tolist(), "grad finite", bool(pt.isfinite(g).all()),
              "gnorm", float(g.norm()), "theta finite", bool(pt.isfinite(S._theta).all()))
if mode == "small":   # for compute-sanitizer: multi-tile per CTA with PSPDE_MAX_GRID=2
    S = mk(100, 64 * 5 + 7, T=0.03); eng = S._get_engine(); theta = S._theta.detach()
    eng.forward(theta, None, Call(offset=0))
    w = pt.randn(eng.K_local, device="cuda"); gr = pt.empty(eng.n_theta, device="cuda")
    eng.backward_detached(theta, w, None, Call(offset=0), gr); pt.cuda.synchronize()
    print("small ok", float(eng.Y_N.sum()), float(gr.norm()))
