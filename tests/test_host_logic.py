"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the Solver's flat
parameter buffer, sharding, the loss/cotangent formulas, problem reference solutions, and a world-size-2 gloo run."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch as pt

from conftest import ROOT, load_golden, relerr
from oracle import manual as man


def test_cabi_library_loads_and_exports_every_declared_symbol():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from pspde import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "pspde.h")).read()
    declared = set(re.findall(r"\b(pspde_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    bound = _lib.bind(_lib.LIB_PATH)                     # no compute calls without a GPU
    assert bound.pspde_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.pspde_cfg) == 120      # static_assert-ed in csrc/api_common.h


def test_product_path_fails_loudly_without_cuda():
    import pspde
    if pt.cuda.is_available():
        pytest.skip("CUDA present")
    prob = pspde.LLGC(d=4, T=0.5, device="cpu")
    with pytest.raises(Exception):
        pspde.Solver("x", prob, K=8, delta_t=0.05, time_approx="inner", verbose=False)       # default device = cuda
    S = pspde.Solver("x", prob, K=8, delta_t=0.05, time_approx="inner", detach_forward=True, verbose=False,
                     device="cpu")
    with pytest.raises(RuntimeError, match="CUDA|no CPU fallback"):
        S.train()
    # same for the diffusion-loss solver; options off the fused path are refused up front
    heat = pspde.HeatEquation(d=4, T=1, device="cpu")
    G = pspde.GeneralSolver(heat, "x", K=8, N=2, delta_t=1e-3, L=1, verbose=False, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA|no CPU fallback"):
        G.train()
    with pytest.raises(NotImplementedError):
        pspde.GeneralSolver(heat, "x", loss_method="BSDE", device="cpu")
    with pytest.raises(NotImplementedError):
        pspde.GeneralSolver(heat, "x", adaptive_forward_process=True, device="cpu")


def test_flat_parameter_buffer_and_module_adam():
    import pspde
    prob = pspde.LQGC(d=3, T=0.25, device="cpu")
    S = pspde.Solver("x", prob, K=8, delta_t=0.05, lr=0.1, detach_forward=True, verbose=False, device="cpu")
    assert S.N == 5 and len(S.z_n) == 5 and S._theta.numel() == 5 * S.p
    off = 0
    for net in S.z_n:
        for q in net.parameters():
            assert q.data_ptr() == S._theta.data_ptr() + 4 * off
            assert q.grad.data_ptr() == S._theta.grad.data_ptr() + 4 * off
            off += q.numel()
    # same seeds -> bit-equal initial weights to the reference's constructors (oracle restates their draw order)
    from oracle import ref_port as orc
    ref = orc.densenet_init(3, 3, seed=42)
    for a, b in zip(S.z_n[0].parameters(), ref):
        assert pt.equal(a.detach(), b)
    before = S._theta.detach().clone()
    S._theta.grad.fill_(1.0)
    S.optimization_step()                               # every module's own Adam steps through the views
    assert pt.allclose(S._theta.detach(), before - 0.1, atol=1e-6)
    S2 = pspde.Solver("x", prob, K=8, delta_t=0.05, time_approx="inner", verbose=False, device="cpu")
    for a, b in zip(S2.z_n.parameters(), orc.mlp_init(4, 3, seed=123)):
        assert pt.equal(a.detach(), b)


def test_flat_adam_equals_per_module_adam():
    """'outer' mode: the single Adam update over the flat buffer is bit-equal to N per-module torch Adams, shares its
    state with them (views), and hands over cleanly when a learning rate is changed or a module is stepped by hand."""
    import pspde
    prob = pspde.LQGC(d=3, T=0.25, device="cpu")
    mk = lambda: pspde.Solver("x", prob, K=8, delta_t=0.05, lr=0.01, detach_forward=True, verbose=False, device="cpu")
    A, B = mk(), mk()
    B._flat_adam = False                                   # B: the per-module loop
    gen = pt.Generator().manual_seed(0)
    for it in range(5):
        g = pt.randn(A._theta.numel(), generator=gen)
        for S in (A, B):
            S._theta.grad.copy_(g)
            S.optimization_step()
        if it == 2:                                        # from here on one module trains slower: both take the loop
            for S in (A, B):
                S.z_n[1].optim.param_groups[0]['lr'] = 0.003
        assert pt.equal(A._theta.detach(), B._theta.detach())
    assert A._flat_adam and not B._flat_adam
    qa, qb = next(A.z_n[2].parameters()), next(B.z_n[2].parameters())
    sa, sb = A.z_n[2].optim.state[qa], B.z_n[2].optim.state[qb]
    assert float(sa['step']) == float(sb['step']) == 5
    assert pt.equal(sa['exp_avg'], sb['exp_avg']) and pt.equal(sa['exp_avg_sq'], sb['exp_avg_sq'])
    assert sa['exp_avg'].data_ptr() >= A._flat_adam['exp_avg'].data_ptr()          # a view of the flat state
    A.z_n[0].optim = pt.optim.Adam(A.z_n[0].parameters(), lr=0.01)                 # replaced optimizer: state rebuilt
    A._theta.grad.fill_(0.5)
    A.optimization_step()
    assert float(A.z_n[0].optim.state[next(A.z_n[0].parameters())]['step']) == 1
    A.z_n[3].optim = pt.optim.SGD(A.z_n[3].parameters(), lr=0.01)                  # not Adam: per-module loop for all
    A.optimization_step()
    assert not A._flat_adam


def test_rows_buffer_sizing():
    """Single-rollout step: the row buffer is whole tiles, bounded by the cap (default 8 GB) and by a quarter of the free
    memory, and not taken at all below a tenth of the batch (pspde/fused.py::rows_buffer_bytes)."""
    from pspde.fused import ROWS_BUFFER_DEFAULT_GB, rows_buffer_bytes
    GB = 2 ** 30
    cap = ROWS_BUFFER_DEFAULT_GB * GB
    tile = 100 * 272 * 512                                   # C2: N = 100, 272 columns x 128 paths x 4 B
    need = 512 * tile                                        # K = 2^16: 7.1 GB
    assert rows_buffer_bytes(need, 512, cap, 170 * GB) == need                     # fits the default cap: every tile
    part = rows_buffer_bytes(need, 512, 2 * GB, 170 * GB)                          # capped: whole tiles below 2 GB
    assert 0 < part <= 2 * GB and part % tile == 0 and part + tile > 2 * GB
    assert rows_buffer_bytes(need, 512, cap, 8 * GB) == int(0.25 * 8 * GB) // tile * tile
    assert rows_buffer_bytes(need, 512, 0.5 * GB, 170 * GB) == 0                   # < 10 % of the batch
    assert rows_buffer_bytes(need, 512, 0, 170 * GB) == 0                          # PSPDE_FWD_CKPT_MAX_GB=0
    c5 = 8192 * 200 * 272 * 512                                                    # C5: a 228 GB tape
    assert rows_buffer_bytes(c5, 8192, cap, 170 * GB) == 0                         # default: no K x N x d tape at C5
    c3 = 2048 * 200 * 176 * 512                                                    # C3: 36.9 GB -> the 8 GB the cap allows
    assert 0.2 < rows_buffer_bytes(c3, 2048, cap, 170 * GB) / c3 < 0.25 and rows_buffer_bytes(c3, 2048, cap, 170 * GB) <= cap
    kept = rows_buffer_bytes(c5, 8192, 96 * GB, 170 * GB)                          # opt-in: what a quarter of the memory holds
    assert kept % (200 * 272 * 512) == 0 and 0.19 < kept / c5 < 0.21


def test_shard_range_covers_everything():
    from pspde.dist import shard_range
    for K in (1, 7, 200, 65536, 65537):
        for W in (1, 2, 3, 8):
            r = [shard_range(K, i, W) for i in range(W)]
            assert r[0][0] == 0 and r[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


@pytest.mark.parametrize("method", ["log-variance", "moment", "variance", "cross_entropy", "relative_entropy"])
def test_loss_cotangents_match_oracle(method):
    from pspde import losses
    rng = np.random.default_rng(1)
    K = 257
    Y, gX, Zs = rng.standard_normal(K), rng.standard_normal(K), rng.random(K)
    lv, wY, wZ = man.loss_and_weights(method, Y, gX, Zs, True)
    t = lambda a: pt.tensor(a, dtype=pt.float32)
    loss, a, b, c, n_bad = losses.value_and_cotangents(method, t(Y), t(gX), t(Zs), K, True)
    assert n_bad.item() == 0
    _, _, _, wG = man.loss_cotangents_full(method, Y, gX, Zs, True)
    assert relerr(c.numpy(), wG) < 1e-5
    assert abs(loss.item() - lv) < 1e-5 * max(1, abs(lv))
    if a is not None:
        assert relerr(a.numpy(), wY) < 1e-5
    else:
        assert np.all(wY == 0)
    if b is not None:
        assert relerr(b.numpy(), wZ) < 1e-6


def test_nonfinite_trajectories_are_dropped_and_counted():
    from pspde import losses
    rng = np.random.default_rng(2)
    K = 64
    Y, gX, Zs = rng.standard_normal(K), rng.standard_normal(K), rng.random(K)
    Yb = Y.copy(); Yb[[3, 17]] = [np.nan, np.inf]
    keep = np.ones(K, bool); keep[[3, 17]] = False
    t = lambda a: pt.tensor(a, dtype=pt.float32)
    for m in ("log-variance", "moment", "variance", "cross_entropy", "relative_entropy"):
        lv, wY, wZ = man.loss_and_weights(m, Y[keep], gX[keep], Zs[keep], True)
        loss, a, b, c, n_bad = losses.value_and_cotangents(m, t(Yb), t(gX), t(Zs), K, True)
        assert n_bad.item() == 2 and abs(loss.item() - lv) < 1e-5 * max(1, abs(lv))
        for w, ref in ((a, wY), (b, wZ)):
            if w is not None:
                assert bool(pt.isfinite(w).all()) and float(w[3]) == 0 and float(w[17]) == 0
                assert relerr(w.numpy()[keep], ref) < 1e-5


def test_problem_reference_solutions():
    import pspde
    p = pspde.LLGC(d=10, off_diag=0.1, T=1, device="cpu")
    v = p.v_true(pt.zeros(1, 10), 0.0)
    assert abs(float(v) - (-2.5651046)) < 1e-5           # SURVEY.md Appendix B known answer
    p = pspde.LLGC(d=100, off_diag=0, T=1, device="cpu")
    assert np.allclose(p.u_true(pt.zeros(3, 100), 0.25), -np.exp(-0.75))
    dw = pspde.DoubleWell(eta=3, kappa=5, device="cpu")
    dw.compute_reference_solution()
    assert abs(float(dw.v_true(pt.tensor([[-1.0, 0.0]]), 0.0)[0, 0]) - 8.6673) < 1e-3
    lq = pspde.LQGC(d=2, device="cpu")
    assert lq.F.shape == (101, 2, 2) and pt.allclose(lq.F[100], pt.eye(2))
    assert float(lq.v_true(pt.zeros(1, 2), 0.0)) == pytest.approx(float(lq.G[0]))
    he = pspde.HeatEquation(d=5, T=1, device="cpu")
    assert float(he.v_true(pt.ones(1, 5), 0.25)) == pytest.approx(5 + 2 * 0.75 * 5)
    for prob in (p, dw, lq, he, pspde.DoubleWell_multidim(d=4, d_1=2, d_2=2, eta=3, kappa=5, device="cpu")):
        pid, flags, pack = prob.functor_pack()
        assert pack.dtype == pt.float32 and pack.numel() >= 7 * prob.d


def test_problem_functors_match_oracle():
    import pspde
    from oracle import ref_port as orc
    x = pt.randn(6, 5)
    z = pt.randn(6, 5)
    for kind, mine, kw in (("llgc", pspde.LLGC, dict(off_diag=0.2, T=1)), ("lqgc", pspde.LQGC, dict(off_diag=0.2, T=1)),
                           ("dwm", pspde.DoubleWell_multidim, dict(d_1=2, d_2=3, eta=3, kappa=5))):
        a, b = mine(d=5, device="cpu", **kw), orc.make_problem(kind, 5, **kw)
        assert pt.allclose(a.b(x), b.b(x), atol=1e-6) and pt.allclose(a.g(x), b.g(x), atol=1e-5)
        assert pt.allclose(a.h(0.1, x, None, z), b.h(0.1, x, None, z), atol=1e-5)
        assert pt.allclose(a.f(x, 0.1), b.f(x, 0.1), atol=1e-5)


def test_elliptic_and_allen_cahn_functors_match_oracle_and_kernel_spec():
    """Host problem classes of the diffusion-loss rows (problems.py:962-1064, :1175-1218, :1546-1580, :1614-1654) against the
    oracle's closures, and the pspde_elliptic spec they hand to the kernels."""
    import pspde
    from oracle import ref_port as orc
    from pspde import _lib as L
    pt.manual_seed(0)
    for kind, mine, d in (("expsphere", pspde.ExponentialOnSphere, 5), ("expball", pspde.ExponentialOnBallNonlinear, 5),
                          ("expball_sin", pspde.ExponentialOnBallNonlinearSin, 5), ("helmholtz", pspde.Helmholtz, 2),
                          ("committor", pspde.Committor, 4)):
        x = pt.randn(7, d) * 0.6
        y = pt.randn(7)
        a, b = mine(d=d, device="cpu"), orc.make_problem(kind, d)
        assert pt.allclose(a.h(x, y, None), b.h(x, y, None), rtol=1e-6, atol=1e-6)
        assert pt.allclose(a.g(x), b.g(x)) and pt.allclose(a.v_true(x), b.v_true(x), rtol=1e-6)
        assert pt.equal(a.sigma(x), b.B) and float(a.b(x).abs().max()) == 0.0
        spec = a.elliptic_spec()
        assert spec.domain == {"sphere": L.DOMAIN_SPHERE, "square": L.DOMAIN_BOX, "two_spheres": L.DOMAIN_ANNULUS}[a.boundary]
        pid, flags, pack = a.functor_pack()
        assert pid == L.PROBLEM_HEAT and flags == 0 and pt.equal(pack[d:2 * d], pt.diag(b.B))
    a, b = pspde.AllenCahn(d=6, T=0.3, device="cpu"), orc.make_problem("allencahn", 6)
    x, y = pt.randn(7, 6), pt.randn(7)
    assert pt.allclose(a.h(0.1, x, y, None), b.h(0.1, x, y, None)) and pt.allclose(a.f(x), b.f(x))
    assert a.functor_pack()[0] == L.PROBLEM_ALLEN_CAHN and a.T == b.T and pt.equal(a.B, b.B)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "path-space-pde-solver_b200"))
import numpy as np, torch as pt, torch.distributed as td
from pspde import losses
from pspde.dist import shard_range, world, all_reduce_sum_
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, W = world()
K = 101
rng = np.random.default_rng(0)
Y, gX, Zs = (pt.tensor(rng.standard_normal(K), dtype=pt.float32) for _ in range(3))
lo, hi = shard_range(K, rank, W)
ok = True
for m in ("log-variance", "moment", "variance", "cross_entropy", "relative_entropy"):
    full = losses.value_and_cotangents.__wrapped__(m, Y, gX, Zs, K) if hasattr(losses.value_and_cotangents, "__wrapped__") else None
    loss, wY, wZ, wG, n_bad = losses.value_and_cotangents(m, Y[lo:hi], gX[lo:hi], Zs[lo:hi], K)
    # single-process reference computed without any process group semantics: emulate by gathering
    parts = [None, None]
    td.all_gather_object(parts, (None if wY is None else wY.numpy(), None if wZ is None else wZ.numpy(), loss.item()))
    if rank == 0:
        from oracle import manual as man
        lv, oY, oZ = man.loss_and_weights(m, Y.double().numpy(), gX.double().numpy(), Zs.double().numpy(), True)
        ok &= abs(parts[0][2] - lv) < 1e-5 * max(1, abs(lv)) and abs(parts[1][2] - lv) < 1e-5 * max(1, abs(lv))
        if parts[0][0] is not None:
            ok &= np.allclose(np.concatenate([parts[0][0], parts[1][0]]), oY, rtol=2e-5, atol=1e-7)
        if parts[0][1] is not None:
            ok &= np.allclose(np.concatenate([parts[0][1], parts[1][1]]), oZ, rtol=2e-5, atol=1e-9)
g = pt.full((5,), float(rank + 1))
all_reduce_sum_(g)
ok &= bool((g == 3).all())
td.destroy_process_group()
if rank == 0:
    print("GLOO_OK" if ok else "GLOO_FAIL")
'''


def test_world_size_2_gloo_sharded_statistics(tmp_path):
    """N>1 host logic on CPU: contiguous sharding + all_reduce of the loss statistics / gradient over gloo give the
    same loss value and per-path cotangents as one process."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0][0], outs


def test_loss_statistics_match_direct_formulas():
    """pspde.losses.value_and_cotangents (one fused all-reduce of the fp64 sums and the dropped-path count) against the
    reference's formulas (solver.py:164-192) evaluated directly, with and without non-finite trajectories."""
    import torch as pt
    from pspde import losses
    g = pt.Generator().manual_seed(0)
    K = 257
    Y, gX, Zs = pt.randn(K, generator=g), pt.randn(K, generator=g), pt.rand(K, generator=g)
    for bad in ((), (3, 100)):
        Yb = Y.clone()
        for i in bad:
            Yb[i] = float("nan")
        ok = pt.isfinite(Yb)
        D = (Yb - gX).double()[ok]
        Ke = float(ok.sum())
        loss, wY, wZ, wG, n_bad = losses.value_and_cotangents("log-variance", Yb, gX, Zs, K)
        assert n_bad.item() == len(bad)
        assert abs(loss.item() - ((D ** 2).mean() - D.mean() ** 2).item()) < 1e-12
        assert pt.allclose(wY[ok].double(), (D - D.mean()) * 2.0 / Ke, atol=1e-7) and bool((wY[~ok] == 0).all()) and wZ is None
        loss, wY, _, _, _ = losses.value_and_cotangents("moment", Yb, gX, Zs, K)
        assert abs(loss.item() - (D ** 2).mean().item()) < 1e-12
        loss, wY, _, _, _ = losses.value_and_cotangents("variance", Yb, gX, Zs, K)
        assert abs(loss.item() - pt.var(pt.exp(D)).item()) < 1e-10
        loss, _, wZ, wG, _ = losses.value_and_cotangents("relative_entropy", Yb, gX, Zs, K)
        assert abs(loss.item() - (Zs.double() + gX.double())[ok].mean().item()) < 1e-12 and pt.allclose(wZ[ok].double().sum(), pt.tensor(1.0, dtype=pt.float64), atol=1e-5)
        # kernel-provided statistics take the place of the local sums
        st = pt.tensor([D.sum(), (D ** 2).sum(), (Zs.double() + gX.double())[ok].sum(), float(len(bad))], dtype=pt.float64)
        loss2, _, _, _, _ = losses.value_and_cotangents("log-variance", Yb, gX, Zs, K, stats=st)
        assert abs(loss2.item() - ((D ** 2).mean() - D.mean() ** 2).item()) < 1e-12


def test_general_solver_inject_draws_follow_the_reference_break():
    """GeneralSolver 'inject' noise: one randn(K, d) per step until every path has stopped (solver.py:1093-1094 before :1106);
    with N * delta_t > T the reference stops drawing early -- the CPU RNG stream must be consumed identically."""
    import torch as pt
    from pspde.general_solver import GeneralSolver

    class Stub:
        pass
    G = Stub()
    G.K, G.d, G.N, G.delta_t_np = 6, 3, 50, 0.01
    G.problem = Stub()
    G.problem.T = 0.3
    pt.manual_seed(11)
    t0 = pt.rand(G.K, 1) * G.problem.T
    xis = GeneralSolver._draw_increments_cpu(G, t0)
    after = pt.rand(1)
    # replay: the reference's loop
    pt.manual_seed(11)
    t = pt.rand(G.K, 1) * G.problem.T
    dt = pt.tensor(G.delta_t_np)
    stopped, n_draws, ref = pt.zeros(G.K, dtype=pt.bool), 0, []
    for n in range(G.N):
        if int((~stopped).sum()) == 0:
            break
        ref.append(pt.randn(G.K, G.d)); n_draws += 1
        ns = (t.squeeze() + dt) <= G.problem.T
        t = t + dt * (ns & ~stopped).float().unsqueeze(1)
        stopped = stopped | (~ns & ~stopped)
    assert n_draws < G.N and pt.equal(after, pt.rand(1))           # same position in the RNG stream afterwards
    assert pt.equal(xis[:n_draws], pt.stack(ref)) and bool((xis[n_draws:] == 0).all()) and xis.shape == (G.N, G.K, G.d)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: ONE JSON line on stdout (everything else goes to stderr); the reference arm runs on the CPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
