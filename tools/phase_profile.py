"""Per-phase cycle counts (CTA 0) of the kernels of one training iteration via pspde_set_profile_buffer, on a bench.py
workload (default c2):  forward rollout (plain and row-keeping), gradient kernel over the rows the forward kept
(single-rollout step, zeta regenerated in the kernel), and their CUDA-event times."""
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
buf = pt.zeros(32, dtype=pt.int64, device=dev)
ntiles = (eng.K_local + 127) // 128
tiles_cta0 = ntiles // 148 + (1 if ntiles % 148 > 0 else 0)


def profiled(fn, warm=True):
    if warm:
        fn(); pt.cuda.synchronize()
    buf.zero_(); lib.pspde_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); pt.cuda.synchronize()
    lib.pspde_set_profile_buffer(None)
    return e0.elapsed_time(e1), buf.tolist()


fwd_names = ["wait G0 (a0 . B0)", "h1 epilogue", "wait G1", "h2 epilogue", "wait G2", "SDE step (+Z ld, a0 st)", "noise (Philox + Box-Muller)"]
for keep in (False, True):
    ms, c = profiled(lambda: eng.forward(theta, None, Call(offset=0), keep_rows=keep))
    steps = tiles_cta0 * eng.N; tot = sum(c[:7])
    print("forward%s: %.2f ms, CTA0 %d tile-steps (128 paths), %.0f cycles/tile-step" % (" (keeps its rows)" if keep else "", ms, steps, tot / steps))
    for n, v in zip(fwd_names, c):
        print("    %-30s %8.0f cycles/tile-step  %5.1f%%" % (n, v / steps, 100 * v / max(tot, 1)))
if eng.ckpt is not None:
    grad_names = ["epi d2: wait hidden MMA 1", "epi d2: wait lo", "epi d2: -", "epi d2: delta_2 (ld, act', st)",
                  "epi d1: waits", "epi d1: delta_1", "gen: Philox + Box-Muller", "gen: wait zeta tile free", "gen: zeta tile stores",
                  "gen: read-back + barrier", "gen: wait A0 free + st", "lo: wait TMA", "lo: fix-up + lo pass", "lo: flush",
                  "tma: wait free buffer", "tma: issue", "mma: wait zeta' + lo (+ flush)", "mma: dW0 issue", "mma: wait delta_2",
                  "mma: hidden MMA 2 issue", "mma: wait zeta tile", "mma: hidden MMA 1 issue", "mma: wait delta_1", "mma: dW1 issue + commits",
                  "epi d2 detail: ld + h LDS + wait_ld + group barrier", "epi d2 detail: sum + act' + split", "epi d2 detail: tmem st issue + 32 STS",
                  "(unused)"]
    ms, c = profiled(lambda: eng.grad_from_rows(theta, wY, Call(offset=0), grad))
    stages = tiles_cta0 * eng.N * 4
    print("gradient kernel over the kept rows: %.2f ms, %.0f cycles per 32-sample stage" % (ms, ms * 1e-3 * 1.965e9 / stages))
    for n, v in zip(grad_names, c):
        print("    %-40s %8.0f cycles/stage" % (n, v / stages))
ms, _ = profiled(lambda: eng.backward_detached(theta, wY, None, Call(offset=0), grad))
print("two-rollout backward (checkpoint rollout + gradient kernel per wave): %.2f ms" % ms)
