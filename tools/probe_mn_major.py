"""Experiment: does tcgen05.mma kind::tf32 accept an MN-major B operand in the 128-byte-swizzled canonical layout?
(The no-swizzle MN-major form returns zeros)  pspde_tc_selftest variant bit 2, descriptor modes in bits 3-4."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
import torch as pt
from pspde import _lib
lib = _lib.load()
K, N = (int(sys.argv[2]) if len(sys.argv) > 2 else 32), 32
pt.manual_seed(0)
A = pt.randn(128, K, device="cuda"); B = pt.randn(K, N, device="cuda")
ref = A.double() @ B.double()
CASES = ((0, 0), (1, 0), (2, 0), (3, 0), (2, 32), (3, 32), (2, 64), (3, 64), (0, 64))
if len(sys.argv) > 1:       # one case per process: a faulting descriptor poisons the context
    CASES = (CASES[int(sys.argv[1])],)
for mode, extra in CASES:
    variant = 4 | (mode << 3) | extra
    D = pt.full((128, N), float("nan"), device="cuda")
    rc = lib.pspde_tc_selftest(K, N, variant, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(D.data_ptr()), None)
    pt.cuda.synchronize()
    err = float((D.double() - ref).norm() / ref.norm())
    print("mode %d data %s (layout type %d): rc %d  rel err %.3e  |D| %.3f" % (mode, "BASE32B atom read K-major" if extra == 64 else "BASE32B atom" if extra else "SW128 atom", 1 if mode & 2 else 2, rc, err, float(D.abs().max())))
