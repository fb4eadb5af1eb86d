"""Per-phase cycle counts of the tensor-core forward kernel (CTA 0, thread 0) via pspde_set_profile_buffer."""
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
names = ["wait G0 (a0 . B0)", "h1 epilogue", "wait G1", "h2 epilogue", "wait G2", "SDE step (+Z ld, a0 st)", "noise (Philox + Box-Muller)"]
ntiles = (eng.K_local + 127) // 128
tiles_cta0 = ntiles // 148 + (1 if ntiles % 148 > 0 else 0)
eng.forward(theta, None, Call(offset=0)); pt.cuda.synchronize()
buf.zero_(); lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
e0.record(); eng.forward(theta, None, Call(offset=0)); e1.record(); pt.cuda.synchronize()
lib.pspde_set_profile_buffer(None)
c = buf.tolist(); tot = sum(c); steps = tiles_cta0 * eng.N
print("tc fwd: %.2f ms, CTA0 %d tile-steps (128 paths), %.0f cycles/tile-step" % (e0.elapsed_time(e1), steps, tot / steps))
for n, v in zip(names, c):
    print("    %-26s %8.0f cycles/tile-step  %5.1f%%" % (n, v / steps, 100 * v / max(tot, 1)))
