"""Per-phase cycle counts of the checkpointed backward (CTA 0) via pspde_set_profile_buffer: the tensor-core checkpoint
rollout and the tensor-core gradient kernel share the 16-slot buffer, so they are profiled in separate passes by
forcing one of them off the profiled path (PSPDE_GRAD_PATH)."""
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
wY = pt.randn(eng.K_local, device=dev) / eng.K_local
grad = pt.empty(eng.n_theta, device=dev)
buf = pt.zeros(16, dtype=pt.int64, device=dev)
eng.backward_detached(theta, wY, None, Call(offset=0), grad); pt.cuda.synchronize()
buf.zero_(); lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
e0.record(); eng.backward_detached(theta, wY, None, Call(offset=0), grad); e1.record(); pt.cuda.synchronize()
lib.pspde_set_profile_buffer(None)
c = buf.tolist()
items = eng.K_local * eng.N / 64 / 148
print("bwd (ckpt path): %.2f ms; ~%.0f gradient items (64 samples) per CTA" % (e0.elapsed_time(e1), items))
print("  slots (cycles, CTA 0 of every launch; rollout slots 0-6 and gradient slots 0-3 overlap):", c[:8])
for n, v in zip(["wait tensor core / rollout G0 wait", "copy + transpose / h1 epi", "hidden cotangents / G1 wait", "fences + MMA issue / h2 epi"], c[:4]):
    print("    %-40s %8.0f cycles/item" % (n, v / items))
