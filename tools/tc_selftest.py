"""GPU check of the tcgen05 3xTF32 building block against fp64 (run under gpurun)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "path-space-pde-solver_b200")):
    sys.path.insert(0, p)
import torch as pt
from pspde import _lib
lib = _lib.load()
pt.manual_seed(0)
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
shapes = ((8, 16), (32, 144), (104, 176), (168, 112))
if len(sys.argv) > 2:
    shapes = shapes[:int(sys.argv[2])]
for variant in variants:
    for (K, N) in shapes:
        A = pt.randn(128, K, device="cuda")
        B = pt.randn(K, N, device="cuda") * 0.1
        D = pt.full((128, N), float("nan"), device="cuda")
        rc = lib.pspde_tc_selftest(K, N, variant, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                   ctypes.c_void_p(D.data_ptr()), None)
        pt.cuda.synchronize()
        ref = A.double() @ B.double()
        e32 = ((A @ B).double() - ref).norm() / ref.norm()
        err = (D.double() - ref).norm() / ref.norm()
        emax = ((D.double() - ref).abs().max() / ref.abs().max()).item()
        print("variant %d K=%3d N=%3d rc=%d  rel err %.3e (max %.3e)   [torch fp32 matmul: %.3e]" % (variant, K, N, rc, err.item(), emax, e32.item()), flush=True)
