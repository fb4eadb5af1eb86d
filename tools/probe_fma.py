import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
from pspde import _lib
lib = _lib.load(); dev = pt.device("cuda", 0)
sink = pt.zeros(4, device=dev); stream = ctypes.c_void_p(pt.cuda.current_stream(dev).cuda_stream)
for mode, name in ((0, "scalar FFMA"), (1, "packed FFMA2"), (2, "FFMA2 + 2 FFMA interleaved")):
    best = 0
    for _ in range(5):
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record(); fl = lib.pspde_fma_probe_ex(mode, 20000, ctypes.c_void_p(sink.data_ptr()), stream); e1.record(); e1.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print("%-30s %.1f TFLOP/s" % (name, best))
