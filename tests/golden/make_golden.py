"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

For every case the reference's own Solver / GeneralSolver / do_importance_sampling_me is run on CPU
(refload.py) with injected Brownian increments, and theta, xi, X_N, Y_N, Z_sum, loss and dLoss/dtheta are
captured at the reference's own call boundary (solver.py:364 initialize_training_data, :202
gradient_descent).  The same inputs are pushed through oracle/ref_port.py and the two are compared here
(tight fp32 tolerances) before anything is written, so a fixture on disk also certifies the oracle.
"""
import os
import sys

import numpy as np
import torch as pt

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import refload  # noqa: E402
from oracle import ref_port as orc  # noqa: E402

pt.set_num_threads(1)
REF = refload.load_reference("cpu")
RS, RP, RF, RU = REF["solver"], REF["problems"], REF["function_space"], REF["utilities"]


def flat(ts):
    return np.concatenate([t.detach().reshape(-1).numpy() for t in ts]).astype(np.float32)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_ref_problem(kind, d, **kw):
    if kind == "llgc":
        return RP.LLGC(d=d, **kw)
    if kind == "lqgc":
        return RP.LQGC(d=d, **kw)
    if kind == "dwm":
        return RP.DoubleWell_multidim(d=d, **kw)
    if kind == "heat":
        return RP.HeatEquation(d=d, **kw)
    raise ValueError(kind)


def run_hjb_case(tag, kind, d, pkw, K, delta_t, net, time_approx, loss_method, detach_forward,
                 adaptive=True, learn_Y_0=False, y0_init=None, xi_seed=7, store_xi=True, lr=1e-3, u_l2=None):
    """One iteration (L=1) of the reference Solver with injected xi; capture everything.
    u_l2: None, or a dict (possibly empty) of keyword arguments for the problem's compute_reference_solution[_2] -- the
    Solver then runs with u_l2_error_flag=True and its u_L2_loss entry (solver.py:491-494, :515) is stored."""
    problem = make_ref_problem(kind, d, **pkw)
    if u_l2 is not None and kind == "dwm":
        problem.compute_reference_solution(**u_l2)
        problem.compute_reference_solution_2(**u_l2)
    S = RS.Solver(tag, problem, lr=lr, L=1, K=K, delta_t=delta_t, loss_method=loss_method,
                  time_approx=time_approx, learn_Y_0=learn_Y_0, adaptive_forward_process=adaptive,
                  detach_forward=detach_forward, early_stopping_time=None, u_l2_error_flag=u_l2 is not None,
                  verbose=False, seed=42)
    if net == "densenet" and time_approx == "inner":
        S.z_n = RF.DenseNet(d_in=d + 1, d_out=d, lr=lr, seed=42)
    if learn_Y_0 and y0_init is not None:
        S.y_0 = RF.SingleParam(lr=lr, initial=y0_init)
    S.update_Phis()
    N = S.N
    g = pt.Generator().manual_seed(xi_seed)
    xi = pt.randn(K, d, N + 1, generator=g)
    cap = {}

    orig_init = S.initialize_training_data

    def init_injected():
        out = list(orig_init())
        out[-1] = xi.clone()
        return tuple(out)

    orig_gd = S.gradient_descent

    def gd_capture(X, Y, Z_sum, l, additional_loss):
        cap["X"], cap["Y"], cap["Zsum"] = X.detach().clone(), Y.detach().clone(), Z_sum.detach().clone()
        cap["gX"] = problem.g(X).detach().clone()
        for phi in S.Phis:  # freeze the optimiser so that theta stays the initial theta
            for grp in phi.optim.param_groups:
                grp["lr"] = 0.0
        return orig_gd(X, Y, Z_sum, l, additional_loss)

    S.initialize_training_data = init_injected
    S.gradient_descent = gd_capture
    nets = S.z_n if time_approx == "outer" else [S.z_n]
    theta0 = [[q.detach().clone() for q in m.parameters()] for m in nets]
    S.train()
    grads = [[(pt.zeros_like(q) if q.grad is None else q.grad.detach().clone()) for q in m.parameters()]
             for m in nets]
    loss = S.loss_log[0]
    gy0 = S.y_0.Y_0.grad.item() if learn_Y_0 else np.nan
    y0v = float(S.y_0.Y_0.detach()) if learn_Y_0 else 0.0

    # ---- oracle cross-check on identical inputs
    op = orc.make_problem(kind, d, **pkw)
    oparams = [[q.clone() for q in net_] for net_ in theta0]
    oparams_arg = oparams if time_approx == "outer" else oparams[0]
    o = orc.hjb_iteration(op, net, oparams_arg, xi.clone(), delta_t, N, loss_method=loss_method,
                          time_approx=time_approx, adaptive=adaptive, detach_forward=detach_forward,
                          y0=pt.tensor([y0v]) if learn_Y_0 else None)
    errs = dict(X=rel(o["X"], cap["X"]), Y=rel(o["Y"], cap["Y"]), loss=abs(float(o["loss"]) - loss) / abs(loss),
                grad=rel(flat(o["grads"]), flat([q for net_ in grads for q in net_])))
    if "relative_entropy" in loss_method:
        errs["Zsum"] = rel(o["Zsum"], cap["Zsum"])
    if learn_Y_0:
        errs["gy0"] = abs(float(o["grad_y0"]) - gy0) / max(abs(gy0), 1e-12)
    print("%-28s loss=%.7e |grad|=%.6e  oracle-vs-ref rel errs: %s" % (
        tag, loss, np.linalg.norm(flat([q for n_ in grads for q in n_])),
        " ".join("%s=%.1e" % kv for kv in errs.items())))
    assert all(v < 2e-5 for v in errs.values()), errs

    out = dict(kind=kind, d=d, K=K, N=N, delta_t=delta_t, T=float(problem.T), net=net, time_approx=time_approx,
               loss_method=loss_method, detach_forward=detach_forward, adaptive=bool(S.adaptive_forward_process),
               learn_Y_0=learn_Y_0, y0=y0v, grad_y0=gy0, xi_seed=xi_seed,
               pkw_keys=np.array(list(pkw.keys())), pkw_vals=np.array([float(v) for v in pkw.values()]),
               theta=flat([q for n_ in theta0 for q in n_]), grad=flat([q for n_ in grads for q in n_]),
               X_N=cap["X"].numpy(), Y_N=cap["Y"].numpy(), Zsum=cap["Zsum"].numpy(), gX=cap["gX"].numpy(),
               loss=np.float64(loss), xi_sum=np.float64(xi.double().sum()), xi_sq=np.float64((xi.double() ** 2).sum()))
    if kind in ("llgc", "lqgc"):
        out["A"] = problem.A.numpy()
        out["B"] = problem.B.numpy()
    if store_xi:
        out["xi"] = xi.numpy()
    if u_l2 is not None:
        out["u_L2_loss"] = np.float64(S.u_L2_loss[0])
        out["ref_kw_keys"] = np.array(list(u_l2.keys()))
        out["ref_kw_vals"] = np.array([float(v) for v in u_l2.values()])
        print("%-28s u_L2_loss=%.7e" % (tag, S.u_L2_loss[0]))
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)


def run_diffusion_case(tag, d, K, K_boundary, N, delta_t, arch, seed=42, full=True, kind="heat", alpha=(1.0, 1.0, 1.0)):
    """GeneralSolver, HeatEquation (or AllenCahn as in its notebook: T = 0.3, boundary_distance = 7, uniform_square),
    loss 'diffusion' (solver.py:1001-1206), L=1, lr frozen to 0."""
    if kind == "heat":
        problem = RP.HeatEquation(d=d, T=1)
    else:
        problem = RP.AllenCahn(d=d, T=0.3)
        problem.modus = "pt"
        problem.boundary_distance = 7.0
    G = RS.GeneralSolver(problem, tag, seed=seed, delta_t=delta_t, N=N, lr=0.0, L=1, K=K, K_boundary=K_boundary,
                         alpha=list(alpha), loss_method="diffusion", verbose=False, uniform_square=(kind != "heat"))
    G.V = RF.DenseNet(d_in=d + 1, d_out=1, lr=0.0, arch=list(arch), seed=seed)
    theta0 = [q.detach().clone() for q in G.V.parameters()]
    G.train()
    grads = [q.grad.detach().clone() for q in G.V.parameters()]
    loss = G.loss_log[0]
    # replicate the reference's draw order (solver.py:1003, :1045-1046, :1078, :1106)
    pt.manual_seed(seed)
    X0 = orc.sample_ball(K, d, problem.boundary_distance) if kind == "heat" else \
        orc.sample_ball_uniform_square(K, d, problem.boundary_distance)
    t0 = pt.rand(K, 1) * problem.T
    xis = pt.stack([pt.randn(K, d) for _ in range(N)])
    op = orc.make_problem("heat", d, T=1) if kind == "heat" else orc.make_problem("allencahn", d, boundary_distance=7.0)
    o = orc.diffusion_iteration(op, [q.clone() for q in theta0], X0, t0, xis, delta_t, N, K_boundary, alpha)
    errs = dict(loss=abs(float(o["loss"]) - loss) / abs(loss), grad=rel(flat(o["grads"]), flat(grads)),
                kcount=abs(o["K_count"] - G.K_log[0]))
    print("%-28s loss=%.7e |grad|=%.6e K_count=%d  oracle-vs-ref: %s" % (
        tag, loss, np.linalg.norm(flat(grads)), G.K_log[0], " ".join("%s=%.1e" % kv for kv in errs.items())))
    assert errs["loss"] < 1e-6 and errs["grad"] < 1e-5 and errs["kcount"] == 0, errs
    out = dict(kind=kind, alpha=np.array(alpha, dtype=np.float64), T=float(problem.T), d=d, K=K, K_boundary=K_boundary, N=N, delta_t=delta_t, arch=np.array(arch), seed=seed,
               loss=np.float64(loss), K_count=G.K_log[0], grad_norm=np.float64(np.linalg.norm(flat(grads))),
               X_end=o["X"].numpy(), t_end=o["t"].numpy(), Y_end=o["Y"].numpy())
    if full:
        out.update(theta=flat(theta0), grad=flat(grads), X0=X0.numpy(), t0=t0.numpy(), xis=xis.numpy())
    else:  # big net: keep a strided sample of the gradient
        g = flat(grads)
        out.update(grad_sample_idx=np.arange(0, g.size, 97), grad_sample=g[::97])
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)


REF_ELLIPTIC = {"expsphere": "ExponentialOnSphere", "expball": "ExponentialOnBallNonlinear",
                "expball_sin": "ExponentialOnBallNonlinearSin", "helmholtz": "Helmholtz", "committor": "Committor"}


def run_elliptic_case(tag, kind, d, K, K_boundary, N, delta_t, arch, alpha, seed=42):
    """EllipticSolver, loss 'diffusion', Dirichlet boundary term (solver.py:628-809), L=1, lr frozen to 0."""
    problem = getattr(RP, REF_ELLIPTIC[kind])(d=d)
    E = RS.EllipticSolver(problem, tag, seed=seed, delta_t=delta_t, N=N, lr=0.0, L=1, K=K, K_boundary=K_boundary,
                          alpha=list(alpha), loss_method="diffusion", verbose=False)
    E.V = RF.DenseNet(d_in=d, d_out=1, lr=0.0, arch=list(arch), seed=seed)
    with pt.no_grad():                       # non-zero biases (the reference initialises them to 0)
        pt.manual_seed(seed + 1)
        for q in E.V.parameters():
            if q.dim() == 1:
                q.add_(0.05 * pt.randn_like(q))
    theta0 = [q.detach().clone() for q in E.V.parameters()]
    E.train()
    grads = [q.grad.detach().clone() for q in E.V.parameters()]
    loss = E.loss_log[0]
    # replicate the reference's draw order (solver.py:630-631, :646-665, :687-708, :726)
    pt.manual_seed(seed)
    np.random.seed(seed)
    op = orc.make_problem(kind, d)
    Xb, X0, _ = orc.elliptic_draws(op, K, K_boundary, 0)
    o = orc.elliptic_iteration(op, [q.clone() for q in theta0], Xb, X0, None, delta_t, N, alpha)
    xis = pt.stack(o["xis"] + [pt.zeros(K, d)] * (N - len(o["xis"])))      # the reference stops drawing once all paths stopped
    errs = dict(loss=abs(float(o["loss"]) - loss) / abs(loss), grad=rel(flat(o["grads"]), flat(grads)),
                kcount=abs(o["K_count"] - E.K_log[0]), vl2=abs(float(o["V_L2"].mean()) - E.V_L2_log[0]) / abs(E.V_L2_log[0]))
    K = X0.shape[0]                                    # 'two_spheres' keeps only the start points in the annulus
    assert K == E.K
    print("%-28s loss=%.7e |grad|=%.6e K_count=%d/%d V_L2=%.6e oracle-vs-ref: %s" % (
        tag, loss, np.linalg.norm(flat(grads)), E.K_log[0], K * N, E.V_L2_log[0],
        " ".join("%s=%.1e" % kv for kv in errs.items())))
    assert errs["loss"] < 1e-6 and errs["grad"] < 1e-5 and errs["kcount"] == 0 and errs["vl2"] < 1e-5, errs
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), kind=kind, d=d, K=K, K_boundary=K_boundary, N=N,
                        delta_t=delta_t, arch=np.array(arch), alpha=np.array(alpha, dtype=np.float64), seed=seed,
                        loss=np.float64(loss), K_count=E.K_log[0], V_L2=np.float64(E.V_L2_log[0]),
                        theta=flat(theta0), grad=flat(grads), Xb=Xb.numpy(), X0=X0.numpy(), xis=xis.numpy(),
                        X_end=o["X"].numpy(), Y_end=o["Y"].numpy(), stopped=o["stopped"].numpy())


def allen_cahn_cases():
    run_diffusion_case("diff_allencahn_d20", 20, 64, 50, 25, 1e-3, (30, 30), kind="allencahn", alpha=(10.0, 1.0, 1.0))

    def g6():       # the 'Allen-Cahn' notebook: d = 100, K = 200, N = 25, dt = 1e-3, alpha = [10, 1, 1], uniform_square
        prob = RP.AllenCahn(d=100, T=0.3)
        prob.modus = "pt"
        prob.boundary_distance = 7.0
        return RS.GeneralSolver(prob, "G6", seed=42, delta_t=1e-3, N=25, lr=1e-3, L=3, K=200, K_boundary=50,
                                alpha=[10.0, 1.0, 1.0], loss_method="diffusion", verbose=False, uniform_square=True)

    run_loss_log_case("loop_G6", g6, 3)


def elliptic_cases():
    run_elliptic_case("ell_expsin_d10", "expball_sin", 10, 64, 20, 20, 1e-3, (30, 30), (0.1, 1.0))   # 'trajectory length' nb
    run_elliptic_case("ell_expball_d5", "expball", 5, 48, 10, 25, 4e-3, (16, 12), (1.0, 1.0))
    run_elliptic_case("ell_expsphere_d4", "expsphere", 4, 40, 10, 30, 1e-2, (12, 8, 8), (0.5, 2.0))  # all paths exit
    run_elliptic_case("ell_helmholtz_d2", "helmholtz", 2, 64, 20, 30, 5e-3, (20, 20), (1.0, 1.0))    # square domain
    run_elliptic_case("ell_committor_d10", "committor", 10, 80, 20, 50, 1e-3, (30, 30), (1.0, 1.0))  # two spheres (K shrinks)

    def g7():       # 'Committor function' notebook: d = 10, K = 200, N = 50, dt = 1e-3
        return RS.EllipticSolver(RP.Committor(d=10), "G7", seed=42, delta_t=1e-3, N=50, lr=1e-3, L=3, K=200,
                                 K_boundary=50, alpha=[1.0, 1.0], loss_method="diffusion", verbose=False)

    def g5():       # 'Nonlinear toy problem - elliptic with Dirichlet' notebook: d=50, K=200, N=20, dt=1e-3
        return RS.EllipticSolver(RP.ExponentialOnBallNonlinearSin(d=50), "G5", seed=42, delta_t=1e-3, N=20, lr=1e-3,
                                 L=3, K=200, K_boundary=50, alpha=[1.0, 1.0], loss_method="diffusion", verbose=False)

    def g5b():
        return RS.EllipticSolver(RP.Helmholtz(d=2), "G5b", seed=42, delta_t=1e-3, N=20, lr=1e-3, L=3, K=200,
                                 K_boundary=50, alpha=[1.0, 1.0], loss_method="diffusion", verbose=False)

    run_loss_log_case("loop_G5", g5, 3)
    run_loss_log_case("loop_G5b", g5b, 3)
    run_loss_log_case("loop_G7", g7, 3)


def run_is_case(tag, kind, d, pkw, K, solver_dt, is_dt, net):
    """do_importance_sampling_me (utilities.py:287-359) on an untrained control."""
    problem = make_ref_problem(kind, d, **pkw)
    S = RS.Solver(tag, problem, K=8, delta_t=solver_dt, time_approx="inner", u_l2_error_flag=False, verbose=False)
    if net == "densenet":
        S.z_n = RF.DenseNet(d_in=d + 1, d_out=d, lr=1e-3, seed=42)
        S.update_Phis()
    theta0 = [q.detach().clone() for q in S.z_n.parameters()]
    N = int(np.ceil(problem.T / is_dt))
    pt.manual_seed(11)
    state = pt.get_rng_state()
    xis = pt.stack([pt.randn(K, d) for _ in range(N)])
    pt.set_rng_state(state)
    mean, var, relerr = RU.do_importance_sampling_me(problem, S, K, delta_t=is_dt)
    op = orc.make_problem(kind, d, **pkw)
    om, ov, orl = orc.importance_sampling(op, net, theta0, xis, is_dt, solver_dt, N_solver=S.N)
    print("%-28s IS mean=%.6e var=%.6e rel=%.6e  oracle: %.6e %.6e %.6e" % (tag, mean, var, relerr, om, ov, orl))
    assert abs(om - mean) / abs(mean) < 1e-5 and abs(ov - var) / abs(var) < 1e-4
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), kind=kind, d=d, K=K, N=N, is_dt=is_dt, solver_dt=solver_dt,
                        net=net, T=float(problem.T), pkw_keys=np.array(list(pkw.keys())),
                        pkw_vals=np.array([float(v) for v in pkw.values()]), theta=flat(theta0), xis=xis.numpy(),
                        mean=mean, var=var, rel=relerr)


def run_loss_log_case(tag, build, L):
    """Whole-loop pins of SURVEY.md Appendix B (loss_log over L iterations incl. Adam; reference RNG)."""
    S = build()
    S.train()
    g = np.sqrt(sum(float((q.grad ** 2).sum()) for phi in S.Phis for q in phi.parameters() if q.grad is not None)) \
        if hasattr(S, "Phis") else np.sqrt(sum(float((q.grad ** 2).sum()) for q in S.V.parameters()))
    print("%-28s loss_log=%s |grad|=%.7e" % (tag, ["%.7e" % v for v in S.loss_log], g))
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), loss_log=np.array(S.loss_log, dtype=np.float64),
                        grad_norm=np.float64(g), K_log=np.array(getattr(S, "K_log", []), dtype=np.int64),
                        V_L2_log=np.array(getattr(S, "V_L2_log", []), dtype=np.float64))


def main():
    only = sys.argv[1:]
    global run_hjb_case
    if only == ["elliptic"]:
        elliptic_cases()
        return
    if only == ["allencahn"]:
        allen_cahn_cases()
        return
    if only:
        _orig = run_hjb_case
        def run_hjb_case(tag, *a, **k):
            if tag in only: _orig(tag, *a, **k)
    # --- single-iteration, injected-noise fixtures (xi stored in the file)
    run_hjb_case("hjb_llgc_d100_dense_lv", "llgc", 100, dict(T=0.2), 16, 0.01, "densenet", "inner",
                 "log-variance", True)                                   # C2/C5 shape (N=20)
    run_hjb_case("hjb_lqgc_d10_dense_lv", "lqgc", 10, dict(T=1), 32, 0.05, "densenet", "inner",
                 "log-variance", True)                                   # C1 inner
    run_hjb_case("hjb_lqgc_d10_outer_lv", "lqgc", 10, dict(T=0.5), 32, 0.05, "densenet", "outer",
                 "log-variance", True)                                   # C1 reference default 'outer'
    run_hjb_case("hjb_dwm_d50_mlp_lv", "dwm", 50, dict(d_1=15, d_2=35, T=0.1, eta=3, kappa=5), 16, 0.005,
                 "mlp_tanh", "inner", "log-variance", True, lr=0.05)     # C3 (i)
    run_hjb_case("hjb_dwm_d50_mlp_re", "dwm", 50, dict(d_1=15, d_2=35, T=0.1, eta=3, kappa=5), 16, 0.005,
                 "mlp_tanh", "inner", "relative_entropy", False, lr=0.05)  # C3 (ii)
    run_hjb_case("hjb_llgc_d10_dense_re", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "relative_entropy", False)
    run_hjb_case("hjb_lqgc_d10_dense_re", "lqgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "relative_entropy", False)
    run_hjb_case("hjb_llgc_d10_moment_y0", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "moment", True, learn_Y_0=True, y0_init=0.7)
    run_hjb_case("hjb_llgc_d10_nonadaptive_lv", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "log-variance", True, adaptive=False)
    run_hjb_case("hjb_llgc_d10_offdiag_lv", "llgc", 10, dict(T=0.5, off_diag=0.1), 32, 0.05, "densenet", "inner",
                 "log-variance", True)                                   # dense A, B
    run_hjb_case("hjb_llgc_d10_crossent", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "cross_entropy", True)
    run_hjb_case("hjb_llgc_d10_variance", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "variance", True)
    run_hjb_case("hjb_llgc_d1_mlp_lv", "llgc", 1, dict(T=0.5), 40, 0.05, "mlp_tanh", "inner",
                 "log-variance", True)                                   # edge: d=1, ragged K
    # attached forward process (the reference's constructor defaults: detach_forward=False, log-variance)
    run_hjb_case("hjb_lqgc_d10_dense_lv_att", "lqgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "log-variance", False)
    run_hjb_case("hjb_dwm_d50_mlp_lv_att", "dwm", 50, dict(d_1=15, d_2=35, T=0.1, eta=3, kappa=5), 16, 0.005,
                 "mlp_tanh", "inner", "log-variance", False, lr=0.05)
    run_hjb_case("hjb_llgc_d10_offdiag_moment_att", "llgc", 10, dict(T=0.5, off_diag=0.1), 32, 0.05, "densenet",
                 "inner", "moment", False)
    run_hjb_case("hjb_lqgc_d10_outer_ce_att", "lqgc", 10, dict(T=0.5), 32, 0.05, "densenet", "outer",
                 "cross_entropy", False)
    # u_L2 diagnostic (u_l2_error_flag=True, the reference default): solver.py:491-494 with problems.py:51-53 (LLGC),
    # :169-171 (LQGC on its OWN Riccati grid delta_t = 0.025 under a solver step of 0.05), :398-404 / :463-476 (double-well
    # finite-difference tables, including the `i[-1] -= 2` element)
    run_hjb_case("hjb_llgc_d10_dense_lv_ul2", "llgc", 10, dict(T=0.5), 32, 0.05, "densenet", "inner",
                 "log-variance", True, u_l2={})
    run_hjb_case("hjb_lqgc_d10_dense_lv_ul2", "lqgc", 10, dict(T=1, delta_t=0.025), 32, 0.05, "densenet", "inner",
                 "log-variance", True, u_l2={})
    run_hjb_case("hjb_dwm_d4_mlp_lv_ul2", "dwm", 4, dict(d_1=2, d_2=2, T=0.3, eta=3, kappa=5), 32, 0.005,
                 "mlp_tanh", "inner", "log-variance", True, lr=0.05, u_l2=dict(delta_t=0.005, nx=400))
    run_hjb_case("hjb_dwm_d4_mlp_re_ul2", "dwm", 4, dict(d_1=2, d_2=2, T=0.3, eta=3, kappa=5), 32, 0.005,
                 "mlp_tanh", "inner", "relative_entropy", False, lr=0.05, u_l2=dict(delta_t=0.005, nx=400))
    if only:
        return
    run_diffusion_case("diff_heat_d10_small", 10, 64, 50, 25, 1e-3, (24, 24), full=True)
    run_diffusion_case("diff_heat_d50_w256", 50, 256, 50, 25, 1e-3, (256, 256), full=False)   # C4 / G4
    run_is_case("is_llgc_d10_dense", "llgc", 10, dict(T=0.5), 64, 0.05, 0.01, "densenet")
    run_is_case("is_dwm_d4_mlp", "dwm", 4, dict(d_1=2, d_2=2, T=0.3, eta=3, kappa=5), 64, 0.005, 0.01, "mlp_tanh")

    # --- whole-loop pins (G1, G1b, G2, G3a, G3b, G4)
    kw = dict(early_stopping_time=None, u_l2_error_flag=False, verbose=False)

    def g1():
        return RS.Solver("G1", RP.LQGC(d=10), K=200, L=3, loss_method="log-variance", detach_forward=True, **kw)

    def g1b():
        S = RS.Solver("G1b", RP.LQGC(d=10), K=200, L=3, loss_method="log-variance", detach_forward=True,
                      time_approx="inner", **kw)
        S.z_n = RF.DenseNet(d_in=11, d_out=10, lr=1e-3, seed=42)
        S.update_Phis()
        return S

    def g2():
        S = RS.Solver("G2", RP.LLGC(d=100, off_diag=0, T=1, seed=42), K=256, L=2, delta_t=0.01,
                      loss_method="log-variance", detach_forward=True, time_approx="inner", **kw)
        S.z_n = RF.DenseNet(d_in=101, d_out=100, lr=1e-3, seed=42)
        S.update_Phis()
        return S

    def g3(loss, detach):
        return lambda: RS.Solver("G3", RP.DoubleWell_multidim(d=50, d_1=15, d_2=35, T=1, eta=3, kappa=5), K=256,
                                 L=2, lr=0.05, delta_t=0.005, loss_method=loss, detach_forward=detach,
                                 time_approx="inner", **kw)

    def g4():
        G = RS.GeneralSolver(RP.HeatEquation(d=50, T=1), "G4", seed=42, delta_t=1e-3, N=25, lr=1e-3, L=2, K=256,
                             K_boundary=50, alpha=[1.0, 1.0, 1.0], loss_method="diffusion", verbose=False)
        G.V = RF.DenseNet(d_in=51, d_out=1, lr=1e-3, arch=[256, 256], seed=42)
        return G

    run_loss_log_case("loop_G1", g1, 3)
    run_loss_log_case("loop_G1b", g1b, 3)
    run_loss_log_case("loop_G2", g2, 2)
    run_loss_log_case("loop_G3a", g3("log-variance", True), 2)
    run_loss_log_case("loop_G3b", g3("relative_entropy", False), 2)
    run_loss_log_case("loop_G4", g4, 2)
    elliptic_cases()
    allen_cahn_cases()


if __name__ == "__main__":
    main()
