"""Issue-to-completion cost of chains of tcgen05.mma kind::tf32 (pspde_mma_probe): SS vs TS operands, M, N, one or several
accumulators.  Decides the shapes of the gradient kernel's hidden-cotangent MMAs."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
from pspde import _lib
lib = _lib.load()
MODES = {0: "SS one acc", 1: "SS two acc", 2: "TS one acc", 3: "SS four acc", 4: "TS precomp", 5: "SS precomp", 6: "TS warp+elect", 7: "SS warp+elect", 8: "2 thr SS|SS", 9: "2 thr SS|TS"}
for n in (16, 96):
    out = pt.zeros(4 * 32, dtype=pt.int64, device="cuda")
    _lib.check(lib, lib.pspde_mma_probe(n, ctypes.c_void_p(out.data_ptr()), None))
    pt.cuda.synchronize()
    o = out.tolist()
    print("n = %d MMAs per chain" % n)
    for c in range(31):
        tot, iss, nn, code = o[4 * c:4 * c + 4]
        if tot == 0:
            break
        mode, M, N = code >> 32, (code >> 16) & 0xffff, code & 0xffff
        print("  %-13s M=%3d N=%3d  %6d cycles to completion (%5.1f / MMA, floor %5.1f), issue loop %5d" % (
            MODES[mode], M, N, tot, tot / nn, max(M, 128) * N / 256.0, iss))
