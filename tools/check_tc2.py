"""grad_tc2_kernel (hidden cotangents + weight gradient on tcgen05, zeta from Philox) against the FP32-FMA recompute backward and
the older tensor-core gradient kernel, Philox noise, several network shapes, ragged K, dead paths, both step forms."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-space-pde-solver_b200")):
    sys.path.insert(0, p)
import torch as pt
import pspde
from pspde.fused import Call


def case(kind, d, arch, K, N, dt=0.02):
    if kind == "dwm":
        prob = pspde.DoubleWell_multidim(d=d, d_1=d // 2, d_2=d - d // 2, T=N * dt, eta=3, kappa=5, device="cuda")
    else:
        prob = pspde.LLGC(d=d, off_diag=0, T=N * dt, seed=42, device="cuda")
    S = pspde.Solver("chk", prob, K=K, L=1, delta_t=dt, time_approx="inner", detach_forward=True, u_l2_error_flag=False,
                     early_stopping_time=None, verbose=False, seed=5)
    if kind != "dwm":
        S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, arch=list(arch), seed=42)
        S.update_Phis()
    eng = S._get_engine()
    theta = S._theta.detach()
    wY = pt.randn(K, device="cuda") / K
    wY[::7] = 0.0                                   # dead paths
    res = {}
    for mode, env in (("simt", dict(PSPDE_BWD_PATH="simt")), ("tc1_waves", dict(PSPDE_GRAD_PATH="tc1")), ("tc2_waves", {}),
                      ("tc2_single", {})):
        for k in ("PSPDE_BWD_PATH", "PSPDE_GRAD_PATH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        g = pt.full((eng.n_theta,), float("nan"), device="cuda")
        c = Call(offset=3)
        kept = eng.forward(theta, None, c, keep_rows=(mode == "tc2_single"))
        if mode == "tc2_single":
            if not kept:
                continue
            eng.grad_from_rows(theta, wY, c, g)
        else:
            eng.backward_detached(theta, wY, None, c, g)
        pt.cuda.synchronize()
        res[mode] = g.double()
    ref = res["simt"]
    out = {k: "%.2e" % float((v - ref).norm() / ref.norm()) for k, v in res.items() if k != "simt"}
    print(kind, d, arch, "K", K, "N", N, out, flush=True)


case("llgc", 100, (30, 30), 128 * 3 + 17, 5)
case("llgc", 10, (30, 30), 300, 4)
case("llgc", 7, (12, 20), 200, 3)
case("dwm", 50, None, 500, 6, dt=0.005)
case("dwm", 6, None, 100, 3, dt=0.005)
case("llgc", 100, (30, 30), 1 << 14, 20, dt=0.01)
