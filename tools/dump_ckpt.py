"""Dump the checkpoint rows the tensor-core forward leaves in the workspace and look for non-finite values."""
import os, sys, ctypes
import numpy as np, torch as pt
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import pspde
from pspde import _lib
from pspde.fused import Call, RolloutEngine
d, K, N = 7, 65, 12
prob = pspde.DoubleWell_multidim(d=d, d_1=d // 3, d_2=d - d // 3, T=1.0, eta=3, kappa=5, device="cuda")
net = pspde.MySequential(d_in=d + 1, d_out=d, lr=1e-3, seed=123).cuda()
theta = pt.cat([q.detach().reshape(-1) for q in net.parameters()]).contiguous()
eng = RolloutEngine(prob, _lib.NET_MLP_TANH, net.net_spec()[1], _lib.TIME_FIRST, K, N, 1.0 / N, seed=5)
wY = pt.randn(K, device="cuda") / K
os.environ["PSPDE_BWD_PATH"] = "ckpt"
eng.workspace.zero_()
g = pt.full((eng.n_theta,), float("nan"), device="cuda")
eng.backward_detached(theta, wY, None, Call(offset=9), g)
pt.cuda.synchronize()
print("grad finite:", bool(pt.isfinite(g).all()), "n_theta", eng.n_theta)
sms = 148
n64 = (K + 63) // 64; grid = min(n64, sms)
al = lambda x: (x + 255) & ~255
stats_b = al(grid * 32)
n128 = (K + 127) // 128; wave = min(n128, sms)
s0 = (d + 2 + 3) // 4 * 4; s0 = (s0 + 7) // 8 * 8
c4 = 2 * (s0 // 4) + 16
items = wave * N * 2; grid_b = min(items, sms)
grad_b = al(grid_b * eng.n_theta * 4)
ck = eng.workspace[stats_b + grad_b: stats_b + grad_b + wave * N * c4 * 128 * 16].view(pt.float32).reshape(wave, N, c4, 128, 4).cpu().numpy()
print("ckpt shape", ck.shape, "finite:", np.isfinite(ck).all())
bad = np.argwhere(~np.isfinite(ck))
print("non-finite count", len(bad), bad[:10])
rows = ck[0, 0].transpose(1, 0, 2).reshape(128, c4 * 4)
np.set_printoptions(linewidth=200, precision=4, suppress=True)
print("row 0:", rows[0]); print("row 64:", rows[64]); print("row 65 (pad):", rows[65][:12])
part = eng.workspace[stats_b: stats_b + grad_b].view(pt.float32).reshape(grid_b, eng.n_theta).cpu().numpy()
print("partials finite per CTA:", np.isfinite(part).all(1)[:24])
