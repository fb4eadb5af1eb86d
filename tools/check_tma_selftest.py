"""TMA (SWIZZLE_128B) + SS-mode tcgen05.mma kind::tf32 building block of the gradient kernel against fp64 numpy.
Prints the relative error of D = T[0:128] . T[rB:rB+N]' and checks the shared-memory image of the first TMA box against
the expected swizzle pattern (16-byte chunk index xor row & 7)."""
import ctypes, os, sys
import numpy as np, torch as pt
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "path-space-pde-solver_b200"))
from pspde import _lib as L
lib = L.load()
rng = np.random.default_rng(0)
ok = True
for R, rB, N in ((256, 128, 112), (272 - 16, 168, 64), (160, 96, 64), (136, 8, 128)):
    T = (rng.standard_normal((R, 128)) * np.exp(rng.uniform(-3, 3, (R, 1)))).astype(np.float32)
    Td = pt.tensor(T).cuda()
    D = pt.full((128, N), float("nan"), device="cuda")
    raw = pt.full((R * 32,), float("nan"), device="cuda")
    rc = lib.pspde_tma_selftest(R, rB, N, Td.data_ptr(), D.data_ptr(), raw.data_ptr(), None)
    if rc != 0:
        print("rc", rc, lib.pspde_last_error()); ok = False; continue
    pt.cuda.synchronize()
    ref = T[:128].astype(np.float64) @ T[rB:rB + N].astype(np.float64).T
    err = np.linalg.norm(D.cpu().numpy() - ref) / np.linalg.norm(ref)
    img = raw.cpu().numpy().reshape(R, 8, 4)                      # [row][physical 16-byte chunk][4 floats]
    exp = np.zeros_like(img)
    for r in range(R):
        for j in range(8):
            exp[r, j ^ (r & 7)] = T[r, 4 * j:4 * j + 4]
    sw_ok = np.array_equal(img, exp)
    print("R=%d rB=%d N=%d: rel err %.2e (fp32-equivalent is ~1e-7), swizzle image %s" % (R, rB, N, err, "ok" if sw_ok else "MISMATCH"))
    if not sw_ok:
        plain = np.array_equal(img.reshape(R, 32), T[:, :32])
        print("   plain (unswizzled) image:", plain, " first rows:", img.reshape(R, 32)[:2, :8], T[:2, :8])
    ok = ok and err < 2e-6 and sw_ok
print("TMA selftest", "PASSED" if ok else "FAILED")
sys.exit(0 if ok else 1)
