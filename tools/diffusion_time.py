"""Times the diffusion-loss kernels at the C4 shape (HeatEquation d=50, DenseNet[256,256], N=25) for a given K."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-space-pde-solver_b200")):
    sys.path.insert(0, p)
import torch as pt
import pspde
from pspde.general_solver import DiffusionEngine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d, N = 50, 25
prob = pspde.HeatEquation(d=d, T=1, device="cuda")
V = pspde.DenseNet(d_in=d + 1, d_out=1, lr=1e-3, arch=[256, 256], seed=42).cuda()
theta = pt.cat([q.detach().reshape(-1) for q in V.parameters()]).contiguous()
eng = DiffusionEngine(prob, V.net_spec()[1], K, N, 1e-3, seed=7)
X0, t0 = eng.sample(1.0, 0)
w = pt.randn(K, device="cuda") / K
grad = pt.empty(eng.n_theta, device="cuda")
dims = V.net_spec()[1]
M = sum(sum(dims[:i + 1]) * dims[i + 1] for i in range(len(dims) - 1))
Md = sum(sum(dims[1:i + 1]) * dims[i + 1] for i in range(1, len(dims) - 1))
tf, tb = [], []
for i in range(reps + 1):
    e = [pt.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); eng.forward(theta, X0, t0, None, i); e[1].record()
    e[2].record(); eng.backward(theta, X0, t0, None, i, -w, w, -w, grad); e[3].record()
    pt.cuda.synchronize()
    if i:
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
tf, tb = statistics.median(tf), statistics.median(tb)
ps = K * N
# algorithmic FLOP per path-step: forward value+tangent 2*2M; backward adds reverse of both 2*2(M + Md) (recompute not counted)
print("K=%d N=%d  fwd %.2f ms (%.2f TFLOP/s)  bwd %.2f ms (%.2f TFLOP/s alg)  step %.3e path-steps/s" % (
    K, N, tf, 4.0 * M * ps / tf / 1e9, tb, 4.0 * (M + Md) * ps / tb / 1e9, ps / ((tf + tb) * 1e-3)))
