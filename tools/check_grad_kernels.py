"""Both gradient kernels (tensor-core, FP32-FMA) against fp64 numpy on synthetic checkpoint rows, several network shapes."""
import os, sys, ctypes
import numpy as np, torch as pt
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
from oracle import manual as man
from pspde import _lib as L
lib = L.load()

def case(kind, d, hid, N=3, n_slots=2, K=200):
    rng = np.random.default_rng(d)
    dims = [d + 1, hid[0], hid[1], d]
    net_id = L.NET_DENSENET if kind == "densenet" else L.NET_MLP_TANH
    cfg = L.make_cfg(K, d, N, 0.01, L.PROBLEM_OU, net_id, dims, L.TIME_FIRST)
    n_theta = lib.pspde_theta_size(ctypes.byref(cfg))
    theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
    net = man.Net(kind, dims, theta.astype(np.float64))
    s0 = (d + 2 + 7) // 8 * 8
    C = 2 * s0 + 64                                   # checkpoint columns [a0 | h1 | h2 | zeta], column-major rows of 128 paths
    ck = np.zeros((n_slots, N, C, 128), np.float32)
    grad = np.zeros(n_theta)
    for slot in range(n_slots):
        live = min(128, K - 128 * slot)
        for n in range(N):
            X = rng.standard_normal((live, d)).astype(np.float32)
            t = np.full((live, 1), 0.01 * n, np.float32)
            zeta = (rng.standard_normal((live, d)) * 0.1).astype(np.float32)
            _, tape = net.forward(np.concatenate([t, X], 1).astype(np.float64))
            grad += net.vjp(tape, zeta.astype(np.float64))
            row = np.zeros((live, C), np.float32)
            row[:, :d], row[:, d:d + 1], row[:, d + 1] = X, t, 1.0
            for l in (0, 1):
                row[:, s0 + 32 * l:s0 + 32 * l + hid[l]] = tape[l][2].astype(np.float32)
                if kind == "mlp_tanh":
                    row[:, s0 + 32 * l + hid[l]] = 1.0
            row[:, s0 + 64:s0 + 64 + d] = zeta
            ck[slot, n, :, :live] = row.T
    th = pt.tensor(theta).cuda(); ckd = pt.tensor(ck).cuda()
    ws = pt.zeros(lib.pspde_workspace_bytes(ctypes.byref(cfg)) + 64 * n_theta * 148, dtype=pt.uint8, device="cuda")
    res = {}
    for name, env in (("simt", dict(PSPDE_GRAD_PATH="simt")), ("tc", dict(PSPDE_GRAD_PATH="tc"))):
        os.environ.update(env)
        out = pt.full((n_theta,), float("nan"), device="cuda")
        rc = lib.pspde_grad_from_ckpt(ctypes.byref(cfg), th.data_ptr(), ckd.data_ptr(), n_slots, s0, out.data_ptr(),
                                      ws.data_ptr(), ws.numel(), None)
        if rc != 0:
            print(name, "rc", rc, lib.pspde_last_error()); continue
        pt.cuda.synchronize()
        o = out.cpu().numpy().astype(np.float64)
        res[name] = np.linalg.norm(o - grad) / np.linalg.norm(grad)
    print(kind, d, hid, {k: "%.2e" % v for k, v in res.items()})

case("densenet", 100, (30, 30), N=4, n_slots=3, K=300)
case("densenet", 10, (30, 30))
case("densenet", 7, (12, 20))
case("mlp_tanh", 6, (30, 30))
case("mlp_tanh", 50, (30, 30))
