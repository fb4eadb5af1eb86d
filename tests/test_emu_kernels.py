"""The CUDA kernel SOURCE (path-space-pde-solver_b200/csrc) executed on the CPU through the fiber emulator
(tests/emu), compared with the reference's golden vectors and the oracle.  These tests check index logic,
shared-memory layout, barrier placement and the host-side planning of the C ABI; the `-m gpu` tests repeat the
same comparisons on the real sm_100a build."""
import ctypes
import os

import numpy as np
import pytest

import emu_harness as H
from conftest import HJB_TAGS, load_golden, relerr
from oracle import manual as man
from oracle import philox as ph
from pspde import _lib as L


@pytest.fixture(scope="module")
def run():
    return H.Runner(H.emu_lib())


def test_emulator_exports_the_whole_abi():
    lib = H.emu_lib()
    assert lib.pspde_abi_version() == L.ABI_VERSION


@pytest.mark.parametrize("tag", HJB_TAGS)
def test_golden_parity(run, tag):
    g = load_golden(tag)
    cfg, pack, x0 = H.cfg_from_golden(g)
    theta, xi = g["theta"].astype(np.float32), np.ascontiguousarray(g["xi"], np.float32)
    K = g["K"]
    if g["detach_forward"]:
        o = run.fwd(cfg, theta, pack, x0, xi, y0=g["y0"] if g["learn_Y_0"] else None)
        assert relerr(o["X"], g["X_N"]) < 1e-5 and relerr(o["Y"], g["Y_N"]) < 1e-5 and relerr(o["gX"], g["gX"]) < 1e-5
        D = o["Y"].astype(np.float64) - o["gX"]
        np.testing.assert_allclose(o["stats"][:2], [D.sum(), (D ** 2).sum()], rtol=1e-12)
        assert o["stats"][3] == 0
        loss, wY, wZ = man.loss_and_weights(g["loss_method"], o["Y"].astype(np.float64), o["gX"].astype(np.float64),
                                            o["Zsum"].astype(np.float64), g["adaptive"])
        tol = 1e-5 * abs(g["loss"]) + (4 * 6e-8 * float((D ** 2).mean()) if "variance" in g["loss_method"] else 0)
        assert abs(loss - g["loss"]) <= tol
        grad = run.bwd(cfg, theta, pack, x0, wY, wZ, xi)
        assert relerr(grad, g["grad"]) < 1e-5
        if g["learn_Y_0"]:
            assert abs(wY.sum() - g["grad_y0"]) < 1e-4 * abs(g["grad_y0"])
    elif g["loss_method"] == "relative_entropy":
        o = run.attached(cfg, theta, pack, x0, 1.0 / K, xi)
        assert relerr(o["X"], g["X_N"]) < 1e-5 and relerr(o["Zsum"], g["Zsum"]) < 1e-5
        assert abs(o["stats"][2] / K - g["loss"]) < 1e-5 * abs(g["loss"])
        assert relerr(o["grad"], g["grad"]) < 1e-5
    else:   # attached forward process, general loss: forward launch -> cotangents -> adjoint launch
        f = run.fwd(cfg, theta, pack, x0, xi)
        assert relerr(f["X"], g["X_N"]) < 1e-5 and relerr(f["Y"], g["Y_N"]) < 1e-5
        loss, wY, wZ, wG = man.loss_cotangents_full(g["loss_method"], f["Y"].astype(np.float64),
                                                    f["gX"].astype(np.float64), f["Zsum"].astype(np.float64), True)
        assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"]) + 4 * 6e-8 * float(((f["Y"] - f["gX"]).astype(np.float64) ** 2).mean())
        o = run.attached(cfg, theta, pack, x0, 0.0, xi, wY=wY, wZ=wZ, wG=wG)
        assert relerr(o["Y"], f["Y"]) < 1e-6 and relerr(o["X"], f["X"]) < 1e-6
        assert relerr(o["grad"], g["grad"]) < 2e-5


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_schedule_independence(order):
    """A missing barrier shows up as a result that depends on the order threads run in between barriers."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import conftest, numpy as np, emu_harness as H;"
            "from conftest import load_golden;"
            "r = H.Runner(H.emu_lib());"
            "g = load_golden('hjb_dwm_d50_mlp_re'); cfg, pack, x0 = H.cfg_from_golden(g);"
            "o = r.attached(cfg, g['theta'], pack, x0, 1.0 / g['K'], np.ascontiguousarray(g['xi']));"
            "g2 = load_golden('hjb_lqgc_d10_outer_lv'); cfg2, pack2, x02 = H.cfg_from_golden(g2);"
            "w = np.linspace(-1, 1, g2['K']).astype(np.float32);"
            "gr = r.bwd(cfg2, g2['theta'], pack2, x02, w, w, np.ascontiguousarray(g2['xi']));"
            "sys.stdout.write('%%.17g %%.17g' %% (float(np.abs(o['grad']).sum()), float(np.abs(gr).sum())))"
            % os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for o in ("forward", order):
        env = dict(os.environ, PSPDE_EMU_ORDER=o)
        outs.append(subprocess.check_output([sys.executable, "-c", code], env=env).decode())
    assert outs[0] == outs[1]


def _philox_case(run, K, k_offset=0, adaptive=True):
    """multi-tile, multi-CTA, ragged last tile, in-kernel Philox; oracle fed with the restated Philox noise."""
    g = load_golden("hjb_lqgc_d10_dense_lv")
    d, N, dt = g["d"], 6, 0.05
    g = dict(g, N=N, adaptive=adaptive)
    cfg, pack, x0 = H.cfg_from_golden(g, noise=L.NOISE_PHILOX, K=K, k_offset=k_offset, seed=1234567, offset=3)
    theta = g["theta"].astype(np.float32)
    o = run.fwd(cfg, theta, pack, x0)
    rng = np.random.default_rng(5)
    wY, wZ = rng.standard_normal(K), rng.standard_normal(K)
    grad = run.bwd(cfg, theta, pack, x0, wY, wZ)
    return g, o, grad, (wY, wZ), theta


def test_philox_multi_tile_vs_oracle(run):
    K = 64 * 4 * 2 + 17           # 9 tiles over the emulator's 4 "SMs", last tile ragged
    g, o, grad, (wY, wZ), theta = _philox_case(run, K)
    d, N, dt = g["d"], g["N"], g["delta_t"]
    xi = ph.xi_tensor(1234567, 3, 0, K, d, N).astype(np.float64)
    net = man.Net("densenet", [d + 1, 30, 30, d], theta)
    prob = man.Problem("lqgc", d)
    gm, ro = man.grad_mode_a(prob, net, xi, dt, N, np.zeros(d), wY, wZ, True, "first")
    assert relerr(o["X"], ro["X"]) < 1e-5 and relerr(o["Y"], ro["Y"]) < 1e-5 and relerr(o["Zsum"], ro["Zsum"]) < 1e-5
    assert relerr(grad, gm) < 2e-5
    D = o["Y"].astype(np.float64) - o["gX"]
    np.testing.assert_allclose(o["stats"][:3], [D.sum(), (D ** 2).sum(), (o["Zsum"].astype(np.float64) + o["gX"]).sum()],
                               rtol=1e-12)


def test_outer_multi_tile_vs_oracle(run):
    """'outer' mode (one parameter set per step) over several tiles per CTA: the per-step weight staging through the index
    table, the weight-image gradient partials that a CTA revisits for every one of its tiles, the row-split gradient
    blocks of a small network and the several-trajectories-per-warp SDE step (d = 10 -> 4 lanes per trajectory)."""
    g = load_golden("hjb_lqgc_d10_outer_lv")
    d, N, dt = g["d"], 5, 0.05
    K = 64 * 4 * 2 + 17           # 9 tiles over the emulator's 4 "SMs", last tile ragged
    n_par = g["theta"].size // g["N"]
    g = dict(g, N=N)
    theta = g["theta"][:N * n_par].astype(np.float32)
    cfg, pack, x0 = H.cfg_from_golden(g, noise=L.NOISE_PHILOX, K=K, seed=4242, offset=1)
    o = run.fwd(cfg, theta, pack, x0)
    rng = np.random.default_rng(11)
    wY, wZ = rng.standard_normal(K), rng.standard_normal(K)
    grad = run.bwd(cfg, theta, pack, x0, wY, wZ)
    assert grad.shape == (N * n_par,) and np.isfinite(grad).all()
    xi = ph.xi_tensor(4242, 1, 0, K, d, N).astype(np.float64)
    nets = [man.Net("densenet", [d, 30, 30, d], theta[n * n_par:(n + 1) * n_par]) for n in range(N)]
    prob = man.Problem("lqgc", d)
    gm, ro = man.grad_mode_a(prob, nets, xi, dt, N, np.zeros(d), wY, wZ, True, "none")
    assert relerr(o["X"], ro["X"]) < 1e-5 and relerr(o["Y"], ro["Y"]) < 1e-5 and relerr(o["Zsum"], ro["Zsum"]) < 1e-5
    assert relerr(grad, gm) < 2e-5
    # every step's block of the gradient is populated (a flush into the wrong slice would leave zeros / double counts)
    per_step = np.abs(grad.reshape(N, n_par)).sum(axis=1)
    assert (per_step > 0).all()


def test_index_table_cache_survives_many_geometries(run):
    """The weight-image index table is cached per network geometry (32 entries, oldest evicted): 40 different networks in one
    process, then the first one again, all against the numpy oracle."""
    d, K, N, dt = 4, 21, 2, 0.1
    pid, flags, pack = H.problem_pack("lqgc", d, {})
    x0 = np.zeros(d, np.float32)
    rng = np.random.default_rng(0)
    wY = rng.standard_normal(K)
    for h in list(range(3, 43)) + [3]:
        dims = [d + 1, h, h + 1, d]
        cfg = L.make_cfg(K, d, N, np.float32(dt), pid, L.NET_DENSENET, dims, L.TIME_FIRST, problem_flags=flags,
                         noise_mode=L.NOISE_PHILOX, seed=5, offset=0)
        n_theta = run.lib.pspde_theta_size(ctypes.byref(cfg))
        theta = (0.3 * np.random.default_rng(h).standard_normal(n_theta)).astype(np.float32)
        grad = run.bwd(cfg, theta, pack, x0, wY, None)
        net = man.Net("densenet", dims, theta)
        xi = ph.xi_tensor(5, 0, 0, K, d, N).astype(np.float64)
        gm, _ = man.grad_mode_a(man.Problem("lqgc", d), net, xi, dt, N, np.zeros(d), wY, np.zeros(K), True, "first")
        assert relerr(grad, gm) < 2e-5, h


def test_misaligned_workspace_is_rejected(run):
    """The kernels address the workspace with 16-byte vector accesses: a misaligned pointer is an error, not a fault."""
    g = load_golden("hjb_lqgc_d10_dense_lv")
    cfg, pack, x0 = H.cfg_from_golden(dict(g, N=2), noise=L.NOISE_PHILOX, K=9, seed=1)
    theta = g["theta"].astype(np.float32)
    lib = run.lib
    ws = np.zeros(lib.pspde_workspace_bytes(ctypes.byref(cfg)) // 8 + 4, np.float64)
    grad = np.zeros(lib.pspde_theta_size(ctypes.byref(cfg)), np.float32)
    wY = np.ones(9, np.float32)
    base = ws.ctypes.data + (-ws.ctypes.data) % 16
    rc = lib.pspde_rollout_bwd_detached(ctypes.byref(cfg), H.ptr(theta), H.ptr(pack), H.ptr(x0), None, H.ptr(wY), None,
                                        H.ptr(grad), ctypes.c_void_p(base + 4), ws.nbytes - 24, None)
    assert rc == -7 and b"aligned" in lib.pspde_last_error()
    rc = lib.pspde_rollout_bwd_detached(ctypes.byref(cfg), H.ptr(theta), H.ptr(pack), H.ptr(x0), None, H.ptr(wY), None,
                                        H.ptr(grad), ctypes.c_void_p(base), ws.nbytes - 24, None)
    assert rc == 0, lib.pspde_last_error()


@pytest.mark.parametrize("d", [20, 132])
def test_attached_kernel_wide_state_vs_oracle(run, d):
    """Attached (relative-entropy) kernel against the numpy discrete adjoint at a state wider than 128 columns -- the backward
    sweep reloads its checkpointed rows four at a time only up to d = 128 and row by row beyond -- and at a narrow one."""
    K, N, dt = 70, 3, 0.05          # two tiles, the second ragged
    pid, flags, pack = H.problem_pack("lqgc", d, {})
    dims = [d + 1, 6, 5, d]         # (tanh MLP: a DenseNet's last layer alone would not fit shared memory at d = 132)
    cfg = L.make_cfg(K, d, N, np.float32(dt), pid, L.NET_MLP_TANH, dims, L.TIME_FIRST, problem_flags=flags,
                     noise_mode=L.NOISE_PHILOX, seed=21, offset=4)
    n_theta = run.lib.pspde_theta_size(ctypes.byref(cfg))
    theta = (0.1 * np.random.default_rng(d).standard_normal(n_theta)).astype(np.float32)
    x0 = np.zeros(d, np.float32)
    o = run.attached(cfg, theta, pack, x0, 1.0 / K)
    xi = ph.xi_tensor(21, 4, 0, K, d, N).astype(np.float64)
    gm, ro = man.grad_mode_b(man.Problem("lqgc", d), man.Net("mlp_tanh", dims, theta), xi, dt, N, np.zeros(d), "first")
    assert relerr(o["X"], ro["X"]) < 1e-5 and relerr(o["Zsum"], ro["Zsum"]) < 1e-5
    assert relerr(o["grad"], gm) < 2e-5


@pytest.mark.parametrize("d", [1, 2, 5, 9, 17, 33, 65])
def test_sde_step_lane_groups_vs_oracle(run, d):
    """The SDE step gives each trajectory G = 1, 2, 4, ... 32 lanes depending on d (several trajectories per warp for small
    d): forward states and the detached gradient against the numpy oracle for one d per group size, ragged last tile."""
    K, N, dt = 64 + 37, 3, 0.05
    pid, flags, pack = H.problem_pack("lqgc", d, {})
    dims = [d + 1, 7, 6, d]
    cfg = L.make_cfg(K, d, N, np.float32(dt), pid, L.NET_DENSENET, dims, L.TIME_FIRST, problem_flags=flags,
                     noise_mode=L.NOISE_PHILOX, seed=8, offset=2)
    n_theta = run.lib.pspde_theta_size(ctypes.byref(cfg))
    theta = (0.2 * np.random.default_rng(100 + d).standard_normal(n_theta)).astype(np.float32)
    x0 = np.zeros(d, np.float32)
    o = run.fwd(cfg, theta, pack, x0)
    rng = np.random.default_rng(d)
    wY, wZ = rng.standard_normal(K), rng.standard_normal(K)
    grad = run.bwd(cfg, theta, pack, x0, wY, wZ)
    xi = ph.xi_tensor(8, 2, 0, K, d, N).astype(np.float64)
    gm, ro = man.grad_mode_a(man.Problem("lqgc", d), man.Net("densenet", dims, theta), xi, dt, N, np.zeros(d), wY, wZ, True, "first")
    assert relerr(o["X"], ro["X"]) < 1e-5 and relerr(o["Y"], ro["Y"]) < 1e-5 and relerr(o["Zsum"], ro["Zsum"]) < 1e-5
    assert relerr(grad, gm) < 2e-5


def test_philox_dump_matches_oracle_and_kernel(run):
    lib = run.lib
    K, d, N = 70, 10, 4
    cfg = L.make_cfg(K, d, N, 0.05, L.PROBLEM_OU, L.NET_DENSENET, [d + 1, 30, 30, d], L.TIME_FIRST, k_offset=5,
                     seed=99, offset=2)
    out = np.zeros((N, K, d), np.float32)
    L.check(lib, lib.pspde_philox_dump(ctypes.byref(cfg), H.ptr(out), None))
    ref = ph.xi_tensor(99, 2, 5, K, d, N)[:, :, 1:].transpose(2, 0, 1)
    assert np.abs(out - ref).max() < 1e-5


def test_sharding_invariance(run):
    """Per-path results do not depend on how K is split over ranks (global-index Philox counter)."""
    K = 150
    _, o, grad, (wY, wZ), theta = _philox_case(run, K)
    g = load_golden("hjb_lqgc_d10_dense_lv")
    g = dict(g, N=6)
    parts, grads = [], []
    for lo, hi in ((0, 80), (80, 150)):
        cfg, pack, x0 = H.cfg_from_golden(g, noise=L.NOISE_PHILOX, K=hi - lo, k_offset=lo, seed=1234567, offset=3)
        parts.append(run.fwd(cfg, theta, pack, x0))
        grads.append(run.bwd(cfg, theta, pack, x0, wY[lo:hi], wZ[lo:hi]))
    assert np.array_equal(np.concatenate([p["Y"] for p in parts]), o["Y"])
    assert np.array_equal(np.concatenate([p["X"] for p in parts]), o["X"])
    assert relerr(grads[0] + grads[1], grad) < 1e-6
    assert abs(sum(p["stats"][0] for p in parts) - o["stats"][0]) < 1e-9 * abs(o["stats"][0]) + 1e-12


def test_error_paths(run):
    lib = run.lib
    d = 10
    bad = L.make_cfg(8, d, 4, 0.05, L.PROBLEM_OU, L.NET_DENSENET, [d + 2, 30, 30, d], L.TIME_FIRST)
    assert lib.pspde_theta_size(ctypes.byref(bad)) < 0 and b"geometry" in lib.pspde_last_error()
    heat = L.make_cfg(8, d, 4, 0.05, L.PROBLEM_HEAT, L.NET_DENSENET, [d + 1, 30, 30, d], L.TIME_FIRST)
    assert lib.pspde_workspace_bytes(ctypes.byref(heat)) == 0
    ok = L.make_cfg(8, d, 4, 0.05, L.PROBLEM_OU, L.NET_DENSENET, [d + 1, 30, 30, d], L.TIME_FIRST,
                    noise_mode=L.NOISE_INJECT)
    theta = np.zeros(lib.pspde_theta_size(ctypes.byref(ok)), np.float32)
    pack, x0 = np.zeros(7 * d, np.float32), np.zeros(d, np.float32)
    ws = np.zeros(1 << 16, np.float64)
    rc = lib.pspde_rollout_fwd(ctypes.byref(ok), H.ptr(theta), H.ptr(pack), H.ptr(x0), None, None, None, None, None,
                               None, None, H.ptr(ws), ws.nbytes, None)
    assert rc < 0 and b"xi" in lib.pspde_last_error()          # INJECT without xi
    rc = lib.pspde_rollout_fwd(ctypes.byref(ok), H.ptr(theta), H.ptr(pack), H.ptr(x0), None, H.ptr(ws), None, None,
                               None, None, None, H.ptr(ws), 8, None)
    assert rc < 0 and b"workspace" in lib.pspde_last_error()


def test_blown_up_trajectory_is_inert(run):
    """One trajectory starts at 1e20 and overflows.  Forward: it is counted and excluded from the statistics.
    Backward (weight 0 from the host) and attached adjoint: gradients stay finite and equal the batch without it."""
    g = load_golden("hjb_lqgc_d10_dense_lv")
    d, N, K = g["d"], 5, 70
    g = dict(g, N=N)
    theta = g["theta"].astype(np.float32)
    bad = 37
    x0 = np.zeros((K, d), np.float32)
    x0[bad] = 1e20
    cfg, pack, _ = H.cfg_from_golden(g, noise=L.NOISE_PHILOX, K=K, seed=11, offset=0)
    cfg.x0_per_path = 1
    o = run.fwd(cfg, theta, pack, x0)
    assert o["stats"][3] == 1 and not np.isfinite(o["Y"][bad] - o["gX"][bad])
    keep = np.arange(K) != bad
    D = o["Y"].astype(np.float64)[keep] - o["gX"][keep]
    np.testing.assert_allclose(o["stats"][:2], [D.sum(), (D ** 2).sum()], rtol=1e-12)
    w = np.random.default_rng(0).standard_normal(K).astype(np.float32)
    w[bad] = 0.0
    grad = run.bwd(cfg, theta, pack, x0, w, w)
    assert np.isfinite(grad).all()
    x0b = x0.copy()
    x0b[bad] = 0.0                      # same batch with a harmless trajectory of zero weight in that slot
    grad_ref = run.bwd(cfg, theta, pack, x0b, w, w)
    assert relerr(grad, grad_ref) < 1e-6
    # attached adjoint
    oa = run.attached(cfg, theta, pack, x0, 1.0 / K)
    assert oa["stats"][3] == 1 and np.isfinite(oa["grad"]).all()
    ob = run.attached(cfg, theta, pack, x0b, 1.0 / K)
    # reference batch: the harmless trajectory contributes; remove its contribution by linearity (run it alone)
    cfg1, _, _ = H.cfg_from_golden(g, noise=L.NOISE_PHILOX, K=1, k_offset=bad, seed=11, offset=0)
    cfg1.x0_per_path = 1
    o1 = run.attached(cfg1, theta, pack, x0b[bad:bad + 1].copy(), 1.0 / K)
    assert relerr(oa["grad"], ob["grad"] - o1["grad"]) < 1e-5


@pytest.mark.parametrize("tag", ["is_llgc_d10_dense", "is_dwm_d4_mlp"])
def test_importance_sampling_rollout(run, tag):
    """pspde_importance_sampling vs the reference's do_importance_sampling_me (utilities.py:287-359) on its own noise."""
    import torch as pt
    g = load_golden(tag)
    d, K, N = g["d"], g["K"], g["N"]
    net = L.NET_DENSENET if g["net"] == "densenet" else L.NET_MLP_TANH
    pid, flags, pack = H.problem_pack(g["kind"], d, g["pkw"])
    xis = np.ascontiguousarray(g["xis"], np.float32)                       # (N, K, d)
    cfg = L.make_cfg(K, d, N, np.float32(g["is_dt"]), pid, net, [d + 1, 30, 30, d], L.TIME_FIRST, adaptive=True,
                     problem_flags=flags, noise_mode=L.NOISE_INJECT, xi_strides=(d, 1, K * d))
    sdt = pt.tensor(g["solver_dt"], dtype=pt.float32)
    t_index = np.array([int(pt.ceil((n * g["is_dt"]) / sdt)) for n in range(N)], np.int32)
    x0 = (-np.ones(d) if g["kind"] == "dwm" else np.zeros(d)).astype(np.float32)
    Y, gX, F = (np.zeros(K, np.float32) for _ in range(3))
    X = np.zeros((K, d), np.float32)
    ws = np.zeros(1 << 14, np.float64)
    lib = run.lib
    rc = lib.pspde_importance_sampling(ctypes.byref(cfg), H.ptr(g["theta"].astype(np.float32)), H.ptr(pack), H.ptr(x0),
                                       H.ptr(xis), H.ptr(t_index), ctypes.c_float(float(sdt)), H.ptr(X), H.ptr(Y),
                                       H.ptr(gX), H.ptr(F), H.ptr(ws), ws.nbytes, None)
    L.check(lib, rc)
    w = np.exp(Y.astype(np.float64) - 2 * F - gX)
    mean, var = w.mean(), w.var(ddof=1)
    assert abs(mean - g["mean"]) < 2e-5 * abs(g["mean"])
    assert abs(var - g["var"]) < 1e-4 * abs(g["var"])
    assert abs(np.sqrt(var) / mean - g["rel"]) < 1e-4 * g["rel"]


@pytest.mark.parametrize("kind", ["llgc", "lqgc", "dwm"])
def test_u_l2_diagnostic(run, kind):
    """u_L2 += sum_j (-Z_j - u*_j(X_{n+1}, n dt))^2 dt (solver.py:491-494) from device tables vs problem.u_true."""
    import torch as pt
    import pspde
    d, K, N, dt = 4, 40, 6, 0.05
    if kind == "llgc":
        prob = pspde.LLGC(d=d, off_diag=0.1, T=N * dt, device="cpu")
    elif kind == "lqgc":
        prob = pspde.LQGC(d=d, T=N * dt, delta_t=0.025, device="cpu")
    else:
        prob = pspde.DoubleWell_multidim(d=d, d_1=2, d_2=2, T=N * dt, eta=3, kappa=5, device="cpu")
        prob.compute_reference_solution(delta_t=0.025, nx=200)
        prob.compute_reference_solution_2(delta_t=0.025, nx=200)
    net = pspde.DenseNet(d_in=d + 1, d_out=d, lr=0.0, seed=3)
    theta = np.concatenate([q.detach().reshape(-1).numpy() for q in net.parameters()]).astype(np.float32)
    pid, flags, pack = prob.functor_pack()
    desc = prob.u_true_table(N, dt)
    tab = np.ascontiguousarray(desc["table"].numpy(), np.float32)
    xi = np.random.default_rng(0).standard_normal((K, d, N + 1)).astype(np.float32)
    cfg = L.make_cfg(K, d, N, np.float32(dt), pid, L.NET_DENSENET, [d + 1, 30, 30, d], L.TIME_FIRST, adaptive=True,
                     problem_flags=flags, noise_mode=L.NOISE_INJECT, xi_strides=(d * (N + 1), N + 1, 1))
    uL2 = np.zeros(K, np.float32)
    u = L.pspde_udiag()
    u.mode, u.nx1, u.d1 = desc["mode"], desc.get("nx1", 0), desc.get("d1", 0)
    u.xb, u.dx = desc.get("xb", 0.0), desc.get("dx", 0.0)
    u.table, u.uL2 = tab.ctypes.data, uL2.ctypes.data
    u.quirk_path = K - 1 if desc.get("quirk_last") else -1
    Y = np.zeros(K, np.float32)
    ws = np.zeros(1 << 14, np.float64)
    x0 = prob.X_0.numpy().astype(np.float32)
    lib = run.lib
    rc = lib.pspde_rollout_fwd_diag(ctypes.byref(cfg), H.ptr(theta), H.ptr(np.ascontiguousarray(pack.numpy())),
                                    H.ptr(x0), None, ctypes.c_void_p(xi.ctypes.data + 4), None, H.ptr(Y), None, None,
                                    None, ctypes.byref(u), H.ptr(ws), ws.nbytes, None)
    L.check(lib, rc)
    # direct evaluation with the problem's own torch functors and u_true (the reference's procedure)
    X = prob.X_0.repeat(K, 1)
    ref = pt.zeros(K)
    xit = pt.tensor(xi)
    with pt.no_grad():
        for n in range(N):
            Z = net(pt.cat([pt.ones(K, 1) * n * pt.tensor(dt), X], 1))
            X = X + (prob.b(X) + (prob.B @ (-Z).t()).t()) * dt + (prob.B @ xit[:, :, n + 1].t()).t() * float(np.sqrt(np.float32(dt)))
            ut = pt.tensor(np.asarray(prob.u_true(X, n * dt))).t().float()
            ref += ((-Z - ut) ** 2).sum(1) * dt
    r = ref.numpy()
    if kind == "dwm":      # including the LAST batch element, whose cell the reference moves by `i[-1] -= 2` (quirk_path)
        assert relerr(uL2, r) < 2e-3
    else:
        assert relerr(uL2, r) < 1e-4


# ---------------------------------------------------------------------------------------------- diffusion loss (a12)
def test_diffusion_golden_parity_small():
    """GeneralSolver iteration (solver.py:1062-1064, :1076-1163) on the reference's own draws: loss, K_count,
    end states, Y and the full gradient against the golden vectors generated from the reference."""
    g = load_golden("diff_heat_d10_small")
    run = H.DiffusionRunner(H.emu_lib())
    o = run.iteration(g, g["theta"].astype(np.float32), g["K_boundary"])
    assert o["K_count"] == g["K_count"]
    assert relerr(o["X"], g["X_end"]) < 1e-6 and relerr(o["t"], g["t_end"].reshape(-1)) < 1e-6
    assert relerr(o["Y"], g["Y_end"]) < 1e-5
    assert abs(o["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    assert relerr(o["grad"], g["grad"]) < 1e-5
    assert abs(np.linalg.norm(o["grad"]) - g["grad_norm"]) < 1e-5 * g["grad_norm"]


def test_diffusion_allen_cahn_golden_parity():
    """GeneralSolver on the Allen-Cahn problem (h = y - y^3, uniform_square start points, alpha = [10, 1, 1]) on the
    reference's own draws: loss, K_count, end states, Y and the full gradient against the golden vectors."""
    g = load_golden("diff_allencahn_d20")
    run = H.DiffusionRunner(H.emu_lib())
    o = run.iteration(g, g["theta"].astype(np.float32), int(g["K_boundary"]), alpha=tuple(g["alpha"]), T=float(g["T"]),
                      kind="allencahn")
    assert o["K_count"] == g["K_count"]
    assert relerr(o["X"], g["X_end"]) < 1e-6 and relerr(o["Y"], g["Y_end"]) < 1e-5
    assert abs(o["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    assert relerr(o["grad"], g["grad"]) < 1e-5


def test_diffusion_ragged_and_stopping():
    """K not a multiple of the tile, paths that start close to T (stopped early) and a 3-hidden-layer net, against
    the fp64 restatement (oracle/manual.py::diffusion)."""
    rng = np.random.default_rng(5)
    K, d, N, dt, arch, T = 37, 6, 7, 0.01, (9, 5, 12), 1.0
    dims = [d + 1] + list(arch) + [1]
    n_theta = sum((sum(dims[:i + 1]) + 1) * dims[i + 1] for i in range(len(dims) - 1))
    theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
    X0 = rng.standard_normal((K, d)).astype(np.float32) * 0.5
    t0 = rng.uniform(0.9, 1.0, K).astype(np.float32)          # about half of the paths hit T within N steps
    xis = rng.standard_normal((N, K, d)).astype(np.float32)
    g = dict(K=K, d=d, N=N, delta_t=dt, arch=arch, X0=X0, t0=t0, xis=xis)
    run = H.DiffusionRunner(H.emu_lib())
    o = run.iteration(g, theta, K_boundary=11)
    net = man.Net("densenet", dims, theta.astype(np.float64))
    m = man.diffusion(man.Problem("heat", d), net, X0.astype(np.float64), t0.astype(np.float64),
                      xis.astype(np.float64), dt, N, 11, T=T)
    assert 0 < o["K_count"] < K * N and o["K_count"] == m["K_count"]
    assert relerr(o["X"], m["X"]) < 1e-6 and relerr(o["Y"], m["Y"]) < 1e-5
    assert abs(o["loss"] - m["loss"]) < 1e-5 * abs(m["loss"])
    assert relerr(o["grad"], m["grad"]) < 2e-5


# ---------------------------------------------------------------------------------------------- elliptic (row f4)
@pytest.mark.parametrize("tag", ["ell_expsin_d10", "ell_expball_d5", "ell_expsphere_d4", "ell_helmholtz_d2", "ell_committor_d10"])
def test_elliptic_golden_parity(tag):
    """EllipticSolver iteration (solver.py:646-670, :687-790) on the reference's own draws: loss, K_count, V_L2,
    end states, Y and the full gradient against golden vectors generated from the reference."""
    g = load_golden(tag)
    run = H.EllipticRunner(H.emu_lib())
    import torch as pt
    from oracle import ref_port as orc
    g_fun = lambda Xb: orc.make_problem(str(g["kind"]), int(g["d"])).g(pt.tensor(Xb, dtype=pt.float32)).numpy().astype(np.float64)
    o = run.iteration(g, g["theta"].astype(np.float32), g_fun, alpha=tuple(g["alpha"]))
    assert o["K_count"] == g["K_count"]
    assert relerr(o["X"], g["X_end"]) < 1e-6
    assert relerr(o["Y"], g["Y_end"]) < 1e-5
    assert abs(o["VL2"].astype(np.float64).mean() - g["V_L2"]) < 1e-5 * g["V_L2"]
    assert abs(o["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    assert relerr(o["grad"], g["grad"]) < 1e-5


def test_elliptic_ragged_one_boundary_and_philox():
    """K not a multiple of the tile, a one-sided box (solver.py:755-756), h == 0, a 3-hidden-layer net, against the
    fp64 restatement; then the same with in-kernel Philox increments read back through pspde_philox_dump."""
    rng = np.random.default_rng(11)
    K, d, N, dt, arch = 45, 3, 9, 0.02, (7, 10, 6)
    dims = [d] + list(arch) + [1]
    n_theta = sum((sum(dims[:i + 1]) + 1) * dims[i + 1] for i in range(len(dims) - 1))
    theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
    X0 = rng.uniform(-1, 1, (K, d)).astype(np.float32)
    xis = rng.standard_normal((N, K, d)).astype(np.float32)
    lib = H.emu_lib()
    run = H.EllipticRunner(lib)
    pack = H.heat_pack(d)
    for one_boundary, h_id in ((True, L.H_ZERO), (False, L.H_EXP_NONLINEAR)):
        ell = L.make_elliptic(L.DOMAIN_BOX, x_l=-1.0, x_r=1.0, one_boundary=one_boundary, h_id=h_id, h_param=(0.5, 0, 0))
        mp = man.EllipticProblem("expball", d, alpha=0.5)
        mp.boundary, mp.one_boundary, mp.X_l, mp.X_r = "square", one_boundary, -1.0, 1.0
        if h_id == L.H_ZERO:
            mp.h = lambda x, y: np.zeros(x.shape[0])
            mp.h_y = lambda x, y: np.zeros(x.shape[0])
        for noise in (L.NOISE_INJECT, L.NOISE_PHILOX):
            cfg = H.elliptic_cfg(K, d, N, dt, arch, noise=noise, seed=7, offset=3, k_offset=5)
            xs = xis
            if noise == L.NOISE_PHILOX:
                xs = np.zeros((N, K, d), np.float32)
                L.check(lib, lib.pspde_philox_dump(ctypes.byref(cfg), H.ptr(xs), None))
            f = run.fwd(cfg, ell, theta, pack, X0, xs if noise == L.NOISE_INJECT else None)
            r = f["VE"].astype(np.float64) - f["Y"]
            w = 2 * r / K
            grad = run.bwd(cfg, ell, theta, pack, X0, xs if noise == L.NOISE_INJECT else None, -w, w, -w)
            net = man.Net("densenet", dims, theta.astype(np.float64))
            m = man.elliptic(mp, net, X0[:1].astype(np.float64), X0.astype(np.float64), xs.astype(np.float64), dt, N,
                             alpha=(1.0, 0.0))
            assert 0 < int(f["stats"][1]) < K * N and int(f["stats"][1]) == m["K_count"]
            assert relerr(f["X"], m["X"]) < 1e-6 and relerr(f["Y"], m["Y"]) < 1e-5
            assert relerr(f["VL2"], m["V_L2"]) < 1e-5
            assert relerr(grad, m["grad"]) < 2e-5


def test_elliptic_error_paths():
    lib = H.emu_lib()
    cfg = H.elliptic_cfg(8, 3, 2, 0.01, (4, 4))
    ell = L.make_elliptic(7)
    assert lib.pspde_elliptic_workspace_bytes(ctypes.byref(cfg), ctypes.byref(ell)) == 0
    assert b"domain" in lib.pspde_last_error()
    bad = H.diffusion_cfg(8, 3, 2, 0.01, (4, 4))              # TIME_LAST network on the elliptic entry point
    ell = L.make_elliptic(L.DOMAIN_SPHERE)
    assert lib.pspde_elliptic_workspace_bytes(ctypes.byref(bad), ctypes.byref(ell)) == 0
    assert b"TIME_NONE" in lib.pspde_last_error()


# ---------------------------------------------------------------------------------------------- checkpointed backward
@pytest.mark.parametrize("kind,d,hid", [("densenet", 10, (30, 30)), ("densenet", 7, (12, 20)), ("mlp_tanh", 6, (30, 30))])
def test_grad_from_checkpoint_rows(kind, d, hid):
    """grad_kernel (second half of the checkpointed detached backward) is a pure function of the operand rows
    [a0 | h1 | h2 | zeta] and theta: dtheta = sum_rows J_theta Z(a0)' zeta.  Rows for 2 tile slots x N steps are built
    with the fp64 network (oracle/manual.py) and random cotangents; padding rows are zero."""
    rng = np.random.default_rng(d)
    lib = H.emu_lib()
    N, n_slots, K = 3, 2, 200                       # 200 of the 256 rows are live
    dims = [d + 1, hid[0], hid[1], d]
    net_id = L.NET_DENSENET if kind == "densenet" else L.NET_MLP_TANH
    cfg = L.make_cfg(K, d, N, 0.01, L.PROBLEM_OU, net_id, dims, L.TIME_FIRST)
    n_theta = lib.pspde_theta_size(ctypes.byref(cfg))
    theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
    net = man.Net(kind, dims, theta.astype(np.float64))
    s0 = (d + 2 + 7) // 8 * 8
    C = 2 * s0 + 64                                   # checkpoint columns [a0 | h1 | h2 | zeta], column-major rows of 128 paths
    ck = np.zeros((n_slots, N, C, 128), np.float32)
    grad = np.zeros(n_theta)
    for slot in range(n_slots):
        live = min(128, K - 128 * slot)
        for n in range(N):
            X = rng.standard_normal((live, d)).astype(np.float32)
            t = np.full((live, 1), 0.01 * n, np.float32)
            zeta = (rng.standard_normal((live, d)) * 0.1).astype(np.float32)
            _, tape = net.forward(np.concatenate([t, X], 1).astype(np.float64))
            grad += net.vjp(tape, zeta.astype(np.float64))
            row = np.zeros((live, C), np.float32)
            row[:, :d], row[:, d:d + 1], row[:, d + 1] = X, t, 1.0
            for l in (0, 1):
                h = tape[l][2].astype(np.float32)
                row[:, s0 + 32 * l:s0 + 32 * l + hid[l]] = h
                if kind == "mlp_tanh":
                    row[:, s0 + 32 * l + hid[l]] = 1.0           # the constant-1 column that carries the next bias
            row[:, s0 + 64:s0 + 64 + d] = zeta
            ck[slot, n, :, :live] = row.T
    out = np.full(n_theta, np.nan, np.float32)
    ws = np.zeros(lib.pspde_workspace_bytes(ctypes.byref(cfg)) // 8 + 8 * n_theta * 64, np.float64)
    rc = lib.pspde_grad_from_ckpt(ctypes.byref(cfg), H.ptr(theta), H.ptr(ck), n_slots, s0, H.ptr(out), H.ptr(ws),
                                  ws.nbytes, None)
    L.check(lib, rc)
    assert relerr(out, grad) < 1e-5


def test_elliptic_annulus_variable_batch():
    """'two_spheres' (Committor, solver.py:694-701, :752-753) in d = 3, where an eighth of the start points falls inside
    the inner sphere and is dropped: exit through either sphere, h == 0, sigma = I, against the fp64 restatement."""
    rng = np.random.default_rng(3)
    d, N, dt, arch = 3, 12, 0.02, (9, 7)
    X = rng.standard_normal((90, d))
    X = 2.0 * X / np.linalg.norm(X, axis=1, keepdims=True) * rng.uniform(0, 1, (90, 1)) ** (1 / d)
    X0 = np.ascontiguousarray(X[np.linalg.norm(X, axis=1) > 1.0], np.float32)
    K = X0.shape[0]
    assert 60 < K < 90
    dims = [d] + list(arch) + [1]
    n_theta = sum((sum(dims[:i + 1]) + 1) * dims[i + 1] for i in range(len(dims) - 1))
    theta = (rng.standard_normal(n_theta) * 0.3).astype(np.float32)
    xis = rng.standard_normal((N, K, d)).astype(np.float32)
    lib = H.emu_lib()
    run = H.EllipticRunner(lib)
    pack = H.heat_pack(d)
    pack[d:2 * d] = 1.0
    ell = H.elliptic_spec("committor")
    cfg = H.elliptic_cfg(K, d, N, dt, arch)
    f = run.fwd(cfg, ell, theta, pack, X0, xis)
    r = f["VE"].astype(np.float64) - f["Y"]
    w = 2 * r / K
    grad = run.bwd(cfg, ell, theta, pack, X0, xis, -w, w, -w)
    net = man.Net("densenet", dims, theta.astype(np.float64))
    m = man.elliptic(man.EllipticProblem("committor", d), net, X0[:1].astype(np.float64), X0.astype(np.float64),
                     xis.astype(np.float64), dt, N, alpha=(1.0, 0.0))
    assert 0 < int(f["stats"][1]) < K * N and int(f["stats"][1]) == m["K_count"]
    assert relerr(f["X"], m["X"]) < 1e-6 and relerr(f["Y"], m["Y"]) < 1e-5
    assert relerr(f["VL2"], m["V_L2"]) < 1e-4
    assert relerr(grad, m["grad"]) < 2e-5
