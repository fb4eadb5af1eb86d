import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "path-space-pde-solver_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)
    out = {k: z[k] for k in z.files}
    for k, v in list(out.items()):
        if isinstance(v, np.ndarray) and v.ndim == 0:
            out[k] = v.item()
    if "pkw_keys" in out:
        out["pkw"] = {str(k): (int(v) if float(v).is_integer() and str(k) != "off_diag" else float(v))
                      for k, v in zip(out["pkw_keys"], out["pkw_vals"])}
    return out


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


HJB_TAGS = ["hjb_llgc_d100_dense_lv", "hjb_lqgc_d10_dense_lv", "hjb_lqgc_d10_outer_lv", "hjb_dwm_d50_mlp_lv",
            "hjb_dwm_d50_mlp_re", "hjb_llgc_d10_dense_re", "hjb_lqgc_d10_dense_re", "hjb_llgc_d10_moment_y0",
            "hjb_llgc_d10_nonadaptive_lv", "hjb_llgc_d10_offdiag_lv", "hjb_llgc_d10_crossent",
            "hjb_llgc_d10_variance", "hjb_llgc_d1_mlp_lv", "hjb_lqgc_d10_dense_lv_att", "hjb_dwm_d50_mlp_lv_att",
            "hjb_llgc_d10_offdiag_moment_att", "hjb_lqgc_d10_outer_ce_att"]


@pytest.fixture
def golden():
    return load_golden
