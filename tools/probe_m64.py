"""Experiment: layout of a tcgen05.mma M = 64 (cta_group::1) accumulator in tensor memory.  A[0:64] sits in lanes 0..63
as for M = 128; all 128 lanes of D are dumped and matched against the rows of A[0:64] . B."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
from pspde import _lib
lib = _lib.load()
K, N = 16, 32
pt.manual_seed(0)
A = pt.randn(128, K, device="cuda"); B = pt.randn(K, N, device="cuda")
for variant in (0, 2):
    D = pt.full((128, N), float("nan"), device="cuda")
    rc = lib.pspde_tc_selftest(K, N, variant, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()), None)
    pt.cuda.synchronize()
    ref = (A.double() @ B.double()).float()
    print("variant", variant, "rc", rc)
    m = {}
    for lane in range(128):
        err = (ref - D[lane][None, :]).abs().max(1).values
        r = int(err.argmin())
        m[lane] = r if float(err[r]) < 1e-3 else None
    print("  lane -> row of A.B:", [m[l] for l in range(128)])
