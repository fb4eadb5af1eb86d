"""Microbenchmark + per-role phase profile of the gradient kernels on a synthetic checkpoint buffer (C2 shape by default):
148 tile slots x N steps of operand rows, timed through pspde_grad_from_ckpt."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import numpy as np, torch as pt
from pspde import _lib as L
lib = L.load()
d, hid, N, slots = 100, (30, 30), int(sys.argv[1]) if len(sys.argv) > 1 else 25, 148
dims = [d + 1, hid[0], hid[1], d]
cfg = L.make_cfg(slots * 128, d, N, 0.01, L.PROBLEM_OU, L.NET_DENSENET, dims, L.TIME_FIRST)
n_theta = lib.pspde_theta_size(ctypes.byref(cfg))
s0 = (d + 2 + 7) // 8 * 8; C = 2 * s0 + 64
gen = pt.Generator(device="cuda").manual_seed(0)
ck = pt.randn(slots, N, C, 128, device="cuda", generator=gen).abs_() * 0.3
theta = pt.randn(n_theta, device="cuda", generator=gen) * 0.1
ws = pt.zeros(lib.pspde_workspace_bytes(ctypes.byref(cfg)) + 4 * n_theta * 160, dtype=pt.uint8, device="cuda")
out = pt.empty(n_theta, device="cuda")
names = ["delta: wait TMA", "delta: zeta.W2' (+delta_2)", "delta: barrier + delta_2.W1' + delta_1", "delta: fence + arrive",
         "mma: wait acc free + lo tile", "mma: wait delta rows", "mma: issue + commit", "mma: flush",
         "lo: wait TMA", "lo: lo pass", "lo: flush", "tma: wait free stage", "tma: issue", "tma: flush"]
for path, flush in (("tc", "16"), ("tc", "100000"), ("tc", "4"), ("simt", "16")):
    os.environ["PSPDE_GRAD_PATH"] = path
    os.environ["PSPDE_GRAD_FLUSH_STAGES"] = flush
    call = lambda: lib.pspde_grad_from_ckpt(ctypes.byref(cfg), theta.data_ptr(), ck.data_ptr(), slots, s0, out.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert call() == 0, lib.pspde_last_error()
    pt.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); pt.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    stages = slots * N * 4 / 148
    print("%s (flush every %s stages): %.3f ms for %d samples -> %.0f cycles per 32-sample stage per SM, %.1f TFLOP/s algorithmic (29 960 MAC/sample)"
          % (path, flush, ms, slots * 128 * N, ms * 1e-3 * 1.965e9 / stages, 2 * 29960 * slots * 128 * N / ms / 1e9))
    if path == "tc":
        buf = pt.zeros(16, dtype=pt.int64, device="cuda")
        lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr())); call(); pt.cuda.synchronize(); lib.pspde_set_profile_buffer(None)
        for n, v in zip(names, buf.tolist()):
            print("    %-40s %8.0f cycles/stage" % (n, v / stages))
