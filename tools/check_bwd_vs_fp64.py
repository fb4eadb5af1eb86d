"""Both detached-backward paths (checkpointed tensor-core, FP32-FMA recompute) against the fp64 restatement
(oracle/manual.py::grad_mode_a) on the kernels' own Philox increments, random per-path cotangents."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import numpy as np, torch as pt
import pspde
from pspde import _lib
from pspde.fused import Call, RolloutEngine
from oracle import manual as man

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

for d, K, N in ((100, 20000, 12), (10, 5000, 40)):
    prob = pspde.LLGC(d=d, T=1.0, device="cuda")
    net = pspde.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, seed=42).cuda()
    theta = pt.cat([q.detach().reshape(-1) for q in net.parameters()]).contiguous()
    eng = RolloutEngine(prob, _lib.NET_DENSENET, net.net_spec()[1], _lib.TIME_FIRST, K, N, 1.0 / N, seed=5)
    gen = pt.Generator(device="cuda").manual_seed(K)
    wY = pt.randn(K, device="cuda", generator=gen) / K
    eng.forward(theta, None, Call(offset=9))
    ok = pt.isfinite(eng.Y_N) & pt.isfinite(eng.gX)
    wY = pt.where(ok, wY, pt.zeros_like(wY))
    xi = eng.philox_dump(offset=9)                       # (K, d, N+1), slice n+1 drives step n
    grads = {}
    for path in ("simt", "ckpt", "ckpt4", "ckpt1"):
        os.environ["PSPDE_BWD_PATH"] = path[:4]
        os.environ["PSPDE_GRAD_FLUSH_ITEMS"] = path[4:] or "64"
        g = pt.empty(eng.n_theta, device="cuda")
        eng.backward_detached(theta, wY, None, Call(offset=9), g)
        pt.cuda.synchronize()
        grads[path] = g.cpu().numpy().astype(np.float64)
    mnet = man.Net("densenet", net.net_spec()[1], theta.cpu().numpy().astype(np.float64))
    mp = man.Problem("llgc", d)
    ref, wYh = 0.0, wY.cpu().numpy().astype(np.float64)
    for lo in range(0, K, 1000):                      # the gradient is a sum over paths: bounded host memory
        hi = min(K, lo + 1000)
        r, _ = man.grad_mode_a(mp, mnet, xi[lo:hi].cpu().numpy().astype(np.float64), np.float32(1.0 / N), N,
                               np.zeros(d), wYh[lo:hi], np.zeros(hi - lo))
        ref = ref + r
    print("d=%d K=%d N=%d: vs fp64: simt %.2e, ckpt (flush 64) %.2e, flush 4 %.2e, flush 1 %.2e; non-finite paths %d"
          % (d, K, N, rel(grads["simt"], ref), rel(grads["ckpt"], ref), rel(grads["ckpt4"], ref), rel(grads["ckpt1"], ref), int((~ok).sum())))
