"""Per-phase cycle counts of the rollout kernels (CTA 0) via the pspde_set_profile_buffer debug hook."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
import bench
from pspde import _lib
from pspde.fused import Call
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = pt.device("cuda", 0); pt.cuda.set_device(0)
lib = _lib.load()
S = bench.build_solver(wl, wl["K"], dev); eng = S._get_engine(); theta = S._theta.detach()
buf = pt.zeros(16, dtype=pt.int64, device=dev)
names = ["prologue", "net_forward(rest)", "sde_step", "backward_hidden", "weight_grad(+copy wait)", "L0 work", "L1 work", "L2 work", "L0 barrier wait", "L1 barrier wait", "L2 barrier wait"]
tiles_cta0 = (eng.K_local + 63) // 64 // 148 + (1 if ((eng.K_local + 63) // 64) % 148 > 0 else 0)
for which in ("fwd", "bwd"):
    eng.forward(theta, None, Call(offset=0)); pt.cuda.synchronize()
    buf.zero_(); lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    if which == "fwd":
        eng.forward(theta, None, Call(offset=0))
    else:
        w = (pt.randn(eng.K_local, device=dev) / eng.K_local) * pt.isfinite(eng.Y_N - eng.gX)
        g = pt.empty(eng.n_theta, device=dev); eng.backward_detached(theta, w.contiguous(), None, Call(offset=0), g)
    e1.record(); pt.cuda.synchronize()
    lib.pspde_set_profile_buffer(None)
    c = buf.tolist(); tot = sum(c); steps = tiles_cta0 * eng.N
    print("%s: %.2f ms, CTA0 %d tile-steps, %.0f cycles/tile-step" % (which, e0.elapsed_time(e1), steps, tot / steps))
    for n, v in zip(names, c):
        if v: print("    %-26s %8.0f cycles/tile-step  %5.1f%%" % (n, v / steps, 100 * v / tot))
