"""Why does the C5 training loss go NaN?  Per iteration: loss, #dropped, max |D| over kept paths, |grad|, |theta|."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-space-pde-solver_b200")):
    sys.path.insert(0, p)
import torch as pt
import pspde
d, K = 100, 1 << 20
prob = pspde.LLGC(d=d, off_diag=0, T=1, seed=42, device="cuda")
S = pspde.Solver("c5", prob, K=K, L=1, delta_t=0.005, time_approx="inner", detach_forward=True, u_l2_error_flag=False,
                 early_stopping_time=None, verbose=False, lr=1e-3)
S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, seed=42)
S.update_Phis()
eng = S._get_engine()
for l in range(14):
    S.train_step(l)
    D = (eng.Y_N.double() - eng.gX.double())
    ok = pt.isfinite(D)
    Dk = D[ok]
    big = (Dk.abs() > 1e6).sum().item()
    print("it %2d loss %.6e dropped %d max|D| %.3e #|D|>1e6 %d |grad| %.4e |theta| %.4e finite_theta %s" % (
        l, S.loss_log[-1], S.nonfinite_log[-1], Dk.abs().max().item() if Dk.numel() else float('nan'), big,
        S._theta.grad.norm().item(), S._theta.norm().item(), bool(pt.isfinite(S._theta).all())))
