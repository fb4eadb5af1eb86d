"""Per-phase cycle counts (thread 0 of CTA 0) of the FP32-FMA rollout kernels (rollout_kernel<FWD|BWD>) on a bench.py
workload outside the tensor-core class (default c1), via pspde_set_profile_buffer."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
import bench
from pspde import _lib
from pspde.fused import Call
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
dev = pt.device("cuda", 0); pt.cuda.set_device(0)
lib = _lib.load()
S = bench.build_solver(wl, wl["K"], dev); eng = S._get_engine(); theta = S._theta.detach()
wY = pt.randn(eng.K_local, device=dev) / eng.K_local
grad = pt.empty(eng.n_theta, device=dev)
buf = pt.zeros(32, dtype=pt.int64, device=dev)
names = {0: "prologue (t column, weights of the step, barrier)", 5: "layer 0 own work", 8: "layer 0 barrier", 6: "layer 1 own work",
         9: "layer 1 barrier", 7: "layer 2 own work", 10: "layer 2 barrier", 1: "(after the network)", 2: "SDE step + barrier",
         3: "hidden cotangents", 11: "weight-gradient accumulation (own blocks)", 12: "weight-gradient flush ('outer')",
         4: "barrier after the weight gradient"}


def profiled(fn):
    fn(); pt.cuda.synchronize()
    buf.zero_(); lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); pt.cuda.synchronize()
    lib.pspde_set_profile_buffer(None)
    return e0.elapsed_time(e1), buf.tolist()


for label, fn in (("forward", lambda: eng.forward(theta, None, Call(offset=0))),
                  ("detached backward", lambda: eng.backward_detached(theta, wY, None, Call(offset=0), grad))):
    ms, c = profiled(fn)
    tot = sum(c)
    print("%s: %.3f ms, %d steps, %.0f cycles/step (event time %.0f cycles/step)" % (label, ms, eng.N, tot / eng.N, ms * 1e-3 * 1.965e9 / eng.N))
    for i in (0, 5, 8, 6, 9, 7, 10, 1, 2, 3, 11, 12, 4):
        if c[i]:
            print("    %-52s %8.0f cycles/step  %5.1f%%" % (names[i], c[i] / eng.N, 100 * c[i] / max(tot, 1)))
