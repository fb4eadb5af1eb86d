"""Debug: raw tensor-memory accumulators of the tensor-core gradient kernel (CTA 0, one item) vs numpy A^T B."""
import os, sys, ctypes
import numpy as np, torch as pt
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
from pspde import _lib as L
lib = L.load()
d, hid, N, K = 10, (30, 30), 1, 64
dims = [d + 1, hid[0], hid[1], d]
cfg = L.make_cfg(K, d, N, 0.01, L.PROBLEM_OU, L.NET_DENSENET, dims, L.TIME_FIRST)
n_theta = lib.pspde_theta_size(ctypes.byref(cfg))
rng = np.random.default_rng(0)
theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
s0 = 16; C4 = 2 * (s0 // 4) + 16
rows = rng.standard_normal((128, 4 * C4)).astype(np.float32)
rows[64:] = 0
rows[:, s0:s0 + 64] = np.abs(rows[:, s0:s0 + 64])       # h >= 0 (sqrt)
ck = np.zeros((1, 1, C4, 128, 4), np.float32)
ck[0, 0] = rows.reshape(128, C4, 4).transpose(1, 0, 2)
th = pt.tensor(theta).cuda(); ckd = pt.tensor(ck).cuda()
ws = pt.zeros(lib.pspde_workspace_bytes(ctypes.byref(cfg)) + 64 * n_theta * 148, dtype=pt.uint8, device="cuda")
nB = ((s0 // 4 + 16) * 4 + 15) // 16 * 16
A = rows[:64, :s0 + 64].astype(np.float64)            # [samples][act cols]
Bz = rows[:64, s0 + 64:s0 + 64 + s0].astype(np.float64)   # zeta part of B only (deltas are computed in-kernel)
ref = A.T @ Bz                                        # [act col][zeta col]
np.set_printoptions(linewidth=220, precision=3, suppress=True)
for variant in (0, 4, 5, 8, 9, 2):
    os.environ.update(PSPDE_GRAD_PATH="tc", PSPDE_GRAD_VARIANT=str(variant))
    dump = pt.zeros(2 * 128 * nB + 64, dtype=pt.float32, device="cuda")
    lib.pspde_set_profile_buffer(ctypes.c_void_p(dump.data_ptr()))
    out = pt.full((n_theta,), float("nan"), device="cuda")
    rc = lib.pspde_grad_from_ckpt(ctypes.byref(cfg), th.data_ptr(), ckd.data_ptr(), 1, s0, out.data_ptr(), ws.data_ptr(), ws.numel(), None)
    pt.cuda.synchronize()
    lib.pspde_set_profile_buffer(None)
    D = dump[:2 * 128 * nB].cpu().numpy().reshape(2, 128, nB)
    got = D[0, :s0 + 64, :s0]
    print("variant", variant, "rc", rc, "|D0|", np.abs(D[0]).max(), "|D1|", np.abs(D[1]).max(), "err vs ref", np.linalg.norm(got - ref) / np.linalg.norm(ref),
          "out finite", bool(pt.isfinite(out).all()), "|out|", float(out.abs().max()))
    tail = dump[2 * 128 * nB:2 * 128 * nB + 64].cpu().numpy()
    print("  tH[0:8]", tail[:8], "rows[0,:4]", rows[0, :4], rows[1, :4], " B[0:8]", tail[32:40], "zeta row0", rows[0, s0 + 64:s0 + 68])
    if variant == 0:
        print("ref[:4,:6]\n", ref[:4, :6]); print("got[:4,:6]\n", got[:4, :6])
        print("got.T[:4,:6]\n", D[0, :6, :4])
