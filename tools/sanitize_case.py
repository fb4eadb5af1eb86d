"""Small workload for compute-sanitizer (memcheck / racecheck): the tcgen05 kernels that have no emulator coverage --
rollout_tc_fwd_kernel (plain, DIAG and row-keeping instantiations) and grad_tc_kernel (single-rollout and wave-checkpointed
launches) -- on the C2 network shape at a few tiles, checked against the FMA kernels in the same run; then the FMA kernels in
the forms that use the index table and the RED.128 flushes: 'outer' mode (C1 shape, several tiles), the attached kernel
(relative entropy) and the diffusion kernels (small C4-like case)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-space-pde-solver_b200")):
    sys.path.insert(0, p)
import torch as pt  # noqa: E402
import pspde  # noqa: E402
from pspde.fused import Call  # noqa: E402

d, K, dt = 100, 128 * 3 + 17, 0.05
prob = pspde.LLGC(d=d, off_diag=0, T=0.25, seed=42, device="cuda")
S = pspde.Solver("san", prob, K=K, L=1, delta_t=dt, time_approx="inner", detach_forward=True, u_l2_error_flag=True,
                 early_stopping_time=None, verbose=False)
S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, seed=42)
S.update_Phis()
eng = S._get_engine()
theta = S._theta.detach()
wY = pt.randn(K, device="cuda") / K
grads = {}
for mode in ("single", "waves", "simt"):
    os.environ.pop("PSPDE_BWD_PATH", None)
    if mode == "simt":
        os.environ["PSPDE_BWD_PATH"] = "simt"
    g = pt.empty(eng.n_theta, device="cuda")
    kept = eng.forward(theta, None, Call(offset=3), keep_rows=(mode == "single"))
    if mode == "single" and kept:
        eng.grad_from_rows(theta, wY, Call(offset=3), g)
    else:
        eng.backward_detached(theta, wY, None, Call(offset=3), g)
    pt.cuda.synchronize()
    grads[mode] = g
for mode in ("single", "waves"):
    err = float((grads[mode] - grads["simt"]).norm() / grads["simt"].norm())
    print(mode, "vs FMA backward: rel err %.2e" % err)
    assert err < 1e-5
S.train_step(0)
print("train_step ok, loss %.6e u_L2 %.6e" % (S.loss_log[-1], S.u_L2_loss[-1]))

# ---- FP32-FMA kernels: 'outer' mode over several tiles, attached kernel, diffusion kernels (bench.py workloads, small K)
import bench  # noqa: E402
for name, K2 in (("c1", 64 * 5 + 9), ("c3re", 64 * 3 + 5)):
    wl = dict(bench.WORKLOADS[name])
    if name == "c3re":
        wl["dt"] = 0.02          # N = 50 (explicit Euler on the double well is unstable at 0.1)
    S2 = bench.build_solver(wl, K2, pt.device("cuda", 0), name="san_" + name)
    for it in range(2):
        S2.train_step(it)
    pt.cuda.synchronize()
    print(name, "train_step ok, loss %.6e" % S2.loss_log[-1])
wl = dict(bench.WORKLOADS["c4"])
dev = pt.device("cuda", 0)
G = pspde.GeneralSolver(pspde.HeatEquation(device=dev, **wl["pkw"]), "san_c4", seed=42, delta_t=wl["dt"], N=4, lr=wl["lr"], L=1,
                        K=300, K_boundary=wl["K_boundary"], verbose=False, device=dev)
G.V = pspde.DenseNet(d_in=wl["pkw"]["d"] + 1, d_out=1, lr=wl["lr"], arch=wl["arch"], seed=42)
for it in range(2):
    G.train_step(it)
pt.cuda.synchronize()
print("c4 train_step ok, loss %.6e" % G.loss_log[-1])
