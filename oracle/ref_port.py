"""CPU oracle for the path-space hot path (TEST INFRASTRUCTURE -- not product code).

A from-scratch PyTorch-CPU restatement of the reference algorithm
(lorenzrichter/path-space-PDE-solver) for the one hot path this repository
accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package (``path-space-pde-solver_b200/pspde``) never does.

Parity pin: every function below is checked in ``tests/test_oracle_golden.py``
against golden vectors produced by the UNMODIFIED reference run on CPU in the
build container (``tests/golden/make_golden.py`` + ``tests/golden/refload.py``).
The reference ships no tests or fixtures of its own.

Reference lines restated here
  * networks ............ function_space.py:116-140 (DenseNet), :177-195 (MySequential)
  * problems ............ problems.py:14-65 (LLGC), :118-175 (LQGC),
                          :285-334 (DoubleWell_multidim), :1733-1764 (HeatEquation)
  * HJB rollout ......... solver.py:364-382 (init), :349-356 (Z_n_), :440-494 (loop)
  * losses .............. solver.py:164-192
  * backward ............ solver.py:202-223 (autograd of the loss)
  * diffusion loss ...... solver.py:1040-1064, :1076-1163, :1187
  * elliptic diffusion .. solver.py:628-826 (EllipticSolver.train, loss 'diffusion', sphere / square domains);
                          problems.py:962-1064 (ExponentialOnSphere / ...OnBallNonlinear / ...NonlinearSin),
                          :1614-1654 (Helmholtz)
  * importance sampling . utilities.py:287-359
"""
from types import SimpleNamespace

import numpy as np
import torch as pt


# --------------------------------------------------------------------------- networks
def densenet_init(d_in, d_out, arch=(30, 30), seed=42, dtype=pt.float32):
    """Parameter list [W0, b0, W1, b1, ...] with the reference's init and draw order
    (function_space.py:117-126): W_i ~ 0.1*N(0,1) of shape (sum(dims[:i+1]), dims[i+1]), b_i = 0."""
    pt.manual_seed(seed)
    dims = [d_in] + list(arch) + [d_out]
    params = []
    for i in range(len(dims) - 1):
        params.append((pt.randn(sum(dims[:i + 1]), dims[i + 1]) * 0.1).to(dtype))
        params.append(pt.zeros(dims[i + 1], dtype=dtype))
    return params


def densenet_forward(params, x):
    """Dense-concat MLP, activation relu(.)**2 (function_space.py:133-140)."""
    n_layers = len(params) // 2
    for i in range(n_layers):
        W, b = params[2 * i], params[2 * i + 1]
        pre = x @ W + b
        if i == n_layers - 1:
            return pre
        x = pt.cat([x, pt.relu(pre) ** 2], dim=1)
    return x


def mlp_init(d_in, d_out, seed=123, dtype=pt.float32):
    """[W0, b0, W1, b1, W2, b2] in nn.Linear layout (out, in) for dims [d_in, 30, 30, d_out];
    nn.Linear default init is drawn first, then overwritten by N(0, 0.01^2) weight-then-bias per layer
    (function_space.py:178-188) -- the draw order matters for bit-equal initial weights."""
    pt.manual_seed(seed)
    dims = [d_in, 30, 30, d_out]
    lins = [pt.nn.Linear(dims[i], dims[i + 1]) for i in range(3)]
    for lin in lins:
        pt.nn.init.normal_(lin.weight, 0, 0.01)
        pt.nn.init.normal_(lin.bias, 0, 0.01)
    out = []
    for lin in lins:
        out += [lin.weight.detach().clone().to(dtype), lin.bias.detach().clone().to(dtype)]
    return out


def mlp_forward(params, x):
    """tanh MLP (function_space.py:190-195)."""
    n_layers = len(params) // 2
    for i in range(n_layers):
        x = x @ params[2 * i].t() + params[2 * i + 1]
        if i < n_layers - 1:
            x = pt.tanh(x)
    return x


NET_FORWARD = {"densenet": densenet_forward, "mlp_tanh": mlp_forward}


# --------------------------------------------------------------------------- problems
def make_problem(kind, d, T=None, dtype=pt.float32, **kw):
    """Problem functors b, sigma(B), h, f, g as closures over plain tensors."""
    p = SimpleNamespace(kind=kind, d=d)
    one = pt.ones(d, dtype=dtype)
    if kind in ("llgc", "lqgc"):
        # problems.py:18-27 / :122-139
        seed = kw.get("seed", 42)
        off = kw.get("off_diag", 0)
        pt.manual_seed(seed)
        p.T = 5 if T is None else T
        p.A = (-pt.eye(d) + off * pt.randn(d, d)).to(dtype)
        p.B = (pt.eye(d) + off * pt.randn(d, d)).to(dtype)
        p.X_0 = pt.zeros(d, dtype=dtype)
        p.b = lambda x: (p.A @ x.t()).t()
        if kind == "llgc":
            p.alpha = pt.ones(d, 1, dtype=dtype)
            p.f = lambda x, t: pt.zeros(x.shape[0], dtype=dtype, device=x.device)
            p.h = lambda t, x, y, z: -0.5 * (z ** 2).sum(1)            # problems.py:45-46
            p.g = lambda x: (x @ p.alpha)[:, 0]                         # :48-49
        else:
            p.P = 0.5 * pt.eye(d, dtype=dtype)
            p.Q = 0.5 * pt.eye(d, dtype=dtype)
            p.R = pt.eye(d, dtype=dtype)
            p.f = lambda x, t: (x.t() * (p.P @ x.t())).sum(0)           # :160-161
            p.g = lambda x: (x.t() * (p.R @ x.t())).sum(0)              # :163-164
            p.h = lambda t, x, y, z: -0.5 * (z ** 2).sum(1) - p.f(x, t)  # :166-167
    elif kind == "dwm":
        # problems.py:289-303, 311-334
        d_1 = kw.get("d_1", d)
        d_2 = kw.get("d_2", 0)
        assert d_1 + d_2 == d
        p.T = 1 if T is None else T
        p.eta_ = pt.tensor([kw.get("eta", 1)] * d_1 + [1.0] * d_2, dtype=dtype)
        p.kappa_ = pt.tensor([kw.get("kappa", 1)] * d_1 + [1.0] * d_2, dtype=dtype)
        p.B = pt.eye(d, dtype=dtype)
        p.X_0 = -pt.ones(d, dtype=dtype)
        p.b = lambda x: -(4.0 * p.kappa_ * (x * (x ** 2 - 1.0)))
        p.f = lambda x, t: pt.zeros(x.shape[0], dtype=dtype, device=x.device)
        p.h = lambda t, x, y, z: -0.5 * (z ** 2).sum(1)
        p.g = lambda x: (p.eta_ * (x - 1.0) ** 2).sum(1)
    elif kind == "heat":
        # problems.py:1734-1758 (f is the *terminal* condition here)
        p.T = 1 if T is None else T
        p.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(dtype)
        p.boundary_distance = 1.0
        p.b = lambda x: pt.zeros_like(x)
        p.h = lambda t, x, y, z: pt.zeros(x.shape[0], dtype=dtype)
        p.f = lambda x: (x ** 2).sum(1)
        p.v_true = lambda x, t: (x ** 2).sum(1) + 2 * (p.T - t) * d
    elif kind == "allencahn":
        # problems.py:1175-1218 (modus 'pt'); the notebook sets boundary_distance = 7.0 after construction
        p.T = 0.3 if T is None else T
        p.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(dtype)
        p.boundary_distance = kw.get("boundary_distance", 2.0)
        p.b = lambda x: pt.zeros_like(x)
        p.h = lambda t, x, y, z: y - y ** 3
        p.f = lambda x: 1 / (2 + 2 / 5 * (x ** 2).sum(1))
    elif kind in ("expsphere", "expball", "expball_sin"):
        # problems.py:962-992, :995-1028, :1031-1064 (elliptic, unit ball, Dirichlet data g = exp(alpha |x|^2))
        a = float(kw.get("alpha", 1.0))
        p.alpha = a
        p.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(dtype)
        p.boundary, p.boundary_distance = "sphere", 1.0
        p.b = lambda x: pt.zeros_like(x)
        p.g = lambda x: pt.exp(a * (x ** 2).sum(1))
        p.v_true = p.g
        if kind == "expsphere":
            p.h = lambda x, y, z: -a * y * (a * 4 * (x ** 2).sum(1) + 2 * d)
        elif kind == "expball":
            p.h = lambda x, y, z: -2 * a * y * (a * 2 * (x ** 2).sum(1) + d) + pt.exp(2 * a * (x ** 2).sum(1)) - y ** 2
        else:
            p.h = lambda x, y, z: -2 * a * y * (a * 2 * (x ** 2).sum(1) + d) + pt.sin(pt.exp(2 * a * (x ** 2).sum(1)) - y ** 2)
    elif kind == "committor":
        # problems.py:1546-1580: two concentric spheres a < |x| < c, sigma = I, h = 0, g = 1 on the outer sphere
        a, c = 1.0, 2.0
        p.a, p.c = a, c
        p.B = pt.eye(d, dtype=dtype)
        p.boundary, p.boundary_distance_1, p.boundary_distance_2 = "two_spheres", a, c
        p.b = lambda x: pt.zeros_like(x)
        p.g = lambda x: (pt.sqrt((x ** 2).sum(1)) > a).to(dtype)
        p.h = lambda x, y, z: pt.zeros(x.shape[0], dtype=dtype)
        p.v_true = lambda x: ((a ** 2 - pt.sqrt((x ** 2).sum(1)) ** (2 - d) * a ** d) / (a ** 2 - c ** (2 - d) * a ** d))
    elif kind == "helmholtz":
        # problems.py:1614-1654 (d = 2, square [-1, 1]^2, both sides absorbing)
        assert d == 2
        p.a_1, p.a_2, p.k = 1.0, 4.0, 1.0
        pi = pt.tensor(np.pi)
        p.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(dtype)
        p.boundary, p.one_boundary, p.X_l, p.X_r = "square", False, -1.0, 1.0
        p.b = lambda x: pt.zeros_like(x)
        ss = lambda x: pt.sin(p.a_1 * pi * x[:, 0]) * pt.sin(p.a_2 * pi * x[:, 1])
        p.g = ss
        p.v_true = ss
        p.h = lambda x, y, z: (p.k ** 2 * y + (p.a_1 * pi) ** 2 * ss(x) + (p.a_2 * pi) ** 2 * ss(x) - p.k ** 2 * ss(x))
    else:
        raise ValueError(kind)
    return p


# --------------------------------------------------------------------------- HJB rollout
def control_eval(net, params, X, n, delta_t, time_approx, N):
    """Z_n_ for approx_method='control' (solver.py:349-356)."""
    if time_approx == "outer":
        n = max(0, min(n, N - 1))
        return NET_FORWARD[net](params[n], X)
    t_col = pt.ones(X.shape[0], 1, dtype=X.dtype, device=X.device) * n * delta_t
    return NET_FORWARD[net](params, pt.cat([t_col, X], 1))


def hjb_rollout(problem, net, params, xi, delta_t, N, time_approx="inner", adaptive=True,
                detach_forward=True, want_zsum=True, y0=None, X0=None, store_path=False):
    """N Euler-Maruyama steps of (X, Y, Z_sum) (solver.py:440-489).

    xi has the reference layout (K, d, N+1); slice n+1 drives step n (:472).
    delta_t is a 0-dim tensor so that t_n, *dt and *sqrt(dt) round as in the reference (:39-40)."""
    K = xi.shape[0]
    dtype, dev = xi.dtype, xi.device
    dt = pt.as_tensor(delta_t, dtype=dtype).to(dev)
    sq = pt.sqrt(dt)
    X = (problem.X_0 if X0 is None else X0).repeat(K, 1) if (X0 is None or X0.dim() == 1) else X0
    Y = pt.zeros(K, dtype=dtype, device=dev) if y0 is None else y0.expand(K) + pt.zeros(K, dtype=dtype, device=dev)
    Zsum = pt.zeros(K, dtype=dtype, device=dev)
    path = [X] if store_path else None
    B = problem.B
    for n in range(N):
        Z = control_eval(net, params, X, n, dt, time_approx, N)                     # :449
        c = -Z.t() if adaptive else pt.zeros(problem.d, K, dtype=dtype, device=dev)             # :451-456
        if detach_forward:
            c = c.detach()                                                          # :468-469
        xin = xi[:, :, n + 1]
        X = X + (problem.b(X) + (B @ c).t()) * dt + (B @ xin.t()).t() * sq          # :471-472
        Y = (Y + (-problem.h(dt * n, X, Y, Z) + (Z * c.t()).sum(1)) * dt
             + (Z * xin).sum(1) * sq)                                               # :477-478
        if want_zsum:
            Zsum = Zsum + (0.5 * (Z ** 2).sum(1) + problem.f(X, n * dt)) * dt       # :486
        if store_path:
            path.append(X)
    return X, Y, Zsum, path


def hjb_loss(loss_method, problem, X, Y, Zsum, adaptive=True):
    """solver.py:164-192 (the variants that are live code)."""
    gX = problem.g(X)
    if loss_method == "moment":
        return ((Y - gX) ** 2).mean()
    if loss_method == "log-variance":
        return ((Y - gX) ** 2).mean() - (Y - gX).mean() ** 2
    if loss_method == "variance":
        return pt.var(pt.exp(-gX + Y))
    if loss_method == "relative_entropy":
        return (Zsum + gX).mean()
    if loss_method == "cross_entropy":
        if adaptive:
            return (Y * pt.exp(-gX + Y.detach())).mean()
        return (Y * pt.exp(-gX)).mean()
    raise ValueError(loss_method)


def flat_params(params, time_approx="inner"):
    if time_approx == "outer":
        return [q for net in params for q in net]
    return list(params)


def hjb_iteration(problem, net, params, xi, delta_t, N, loss_method="log-variance", time_approx="inner",
                  adaptive=True, detach_forward=True, y0=None):
    """One training iteration's loss and dLoss/dtheta by autograd (solver.py:433-499, :220-221).

    Returns dict(loss, grads (list, parameter order), X, Y, Zsum, grad_y0)."""
    if loss_method == "relative_entropy":
        adaptive = True                                                             # :61-62
    leaves = flat_params(params, time_approx)
    for q in leaves:
        q.requires_grad_(True)
        q.grad = None
    if y0 is not None:
        y0.requires_grad_(True)
        y0.grad = None
    X, Y, Zsum, _ = hjb_rollout(problem, net, params, xi, delta_t, N, time_approx, adaptive, detach_forward,
                                want_zsum="relative_entropy" in loss_method, y0=y0)
    loss = hjb_loss(loss_method, problem, X, Y, Zsum, adaptive)
    loss.backward()
    grads = [pt.zeros_like(q) if q.grad is None else q.grad.detach().clone() for q in leaves]
    out = dict(loss=loss.detach(), grads=grads, X=X.detach(), Y=Y.detach(), Zsum=Zsum.detach(),
               gX=problem.g(X).detach(), grad_y0=None if y0 is None else y0.grad.detach().clone())
    for q in leaves:
        q.requires_grad_(False)
    return out


# --------------------------------------------------------------------------- diffusion loss (GeneralSolver)
def diffusion_iteration(problem, params, X0, t0, xis, delta_t, N, K_boundary=50, alpha=(1.0, 1.0, 1.0)):
    """One iteration of GeneralSolver.train for loss_method='diffusion', boundary='unbounded',
    non-adaptive, detach_forward=True (solver.py:1062-1064, :1076-1163).

    X0 (K,d), t0 (K,1), xis (N,K,d) are the random draws (:1045-1046, :1078, :1106) supplied by the caller.
    The network sees [X, t] with t as the LAST column (:1079)."""
    dtype = X0.dtype
    dt = pt.as_tensor(delta_t, dtype=dtype)
    sq = pt.sqrt(dt)
    K = X0.shape[0]
    for q in params:
        q.requires_grad_(True)
        q.grad = None
    V = lambda z: densenet_forward(params, z)
    X_T = pt.cat([X0[:K_boundary], problem.T * pt.ones(K_boundary, 1, dtype=dtype)], 1)
    loss = alpha[1] * ((V(X_T).squeeze() - problem.f(X0[:K_boundary])) ** 2).mean()  # :1063-1064
    X = X0.clone().requires_grad_(True)
    t_n = t0.clone()
    X_t = pt.cat([X, t_n], 1)
    Y = V(X_t).squeeze()                                                             # :1081
    stopped = pt.zeros(K, dtype=pt.bool)
    K_count = 0
    for n in range(N):
        if int((~stopped).sum()) == 0:
            break
        Y_ = V(X_t)
        grad_V, = pt.autograd.grad(Y_.squeeze().sum(), X, create_graph=True)         # :1100-1103
        Z = (problem.B.t() @ grad_V.t()).t()                                         # :1104
        xin = xis[n]
        sel = (~stopped).to(dtype).unsqueeze(1)
        X_prop = X + (problem.b(X) * dt + (problem.B @ xin.t()).t() * sq) * sel      # :1116-1117 (c = 0)
        new_sel = (t_n.squeeze(1) + dt) <= problem.T                                 # :1119, :1131
        act = (new_sel & ~stopped)
        actf = act.to(dtype)
        Y = Y + (-problem.h(n * dt, X, Y_.squeeze(), Z) * dt + (Z * xin).sum(1) * sq) * actf   # :1141-1142
        X = X * (1 - actf).unsqueeze(1) + X_prop * actf.unsqueeze(1)                 # :1145-1146
        t_n = t_n + dt * actf.unsqueeze(1)                                           # :1148
        X_t = pt.cat([X, t_n], 1)
        K_count += int(act.sum())                                                    # :1151-1152
        stopped = stopped | (~new_sel & ~stopped)                                    # :1154-1155
    loss = loss + alpha[0] * ((V(X_t).squeeze() - Y) ** 2).mean()                    # :1163
    loss.backward()
    grads = [q.grad.detach().clone() for q in params]
    for q in params:
        q.requires_grad_(False)
    return dict(loss=loss.detach(), grads=grads, K_count=K_count, X=X.detach(), t=t_n.detach(), Y=Y.detach())


# --------------------------------------------------------------------------- elliptic diffusion loss (EllipticSolver)
def elliptic_exit_mask(problem, X, X_prop):
    """new_selection of solver.py:750-760: the sphere test looks at X (before the step), the square test at the
    proposal."""
    if problem.boundary == "sphere":
        return pt.sqrt((X ** 2).sum(1)) < problem.boundary_distance                  # :750-751
    if problem.boundary == "two_spheres":
        r = pt.sqrt((X ** 2).sum(1))
        return (r > problem.boundary_distance_1) & (r < problem.boundary_distance_2)   # :752-753
    if problem.boundary == "square":
        if problem.one_boundary:
            return (X_prop <= problem.X_r).all(1)                                    # :755-756
        return ((X_prop >= problem.X_l) & (X_prop <= problem.X_r)).all(1)            # :757-758
    raise ValueError(problem.boundary)


def elliptic_iteration(problem, params, Xb, X0, xis, delta_t, N, alpha=(1.0, 1.0)):
    """One iteration of EllipticSolver.train for loss_method='diffusion', approx_method='Y', Dirichlet boundary
    term, non-adaptive, detach_forward=True (solver.py:628-790).  Xb (K_boundary, d) are the boundary samples
    (:646-648 / :655-665), X0 (K, d) the interior start points (:687-708), xis (N, K, d) the increments (:726);
    with xis=None they are drawn here, one pt.randn(K, d) per step BEFORE the all-stopped check as the reference
    does (:726-730), so the RNG stream of a whole training loop is reproduced (the draws are returned).
    The value network sees X only.  Also returns the V_L2 diagnostic of :733 (per path)."""
    dtype = X0.dtype
    dt = pt.as_tensor(delta_t, dtype=dtype)
    sq = pt.sqrt(dt)
    K = X0.shape[0]
    for q in params:
        q.requires_grad_(True)
        q.grad = None
    V = lambda z: densenet_forward(params, z)
    loss = alpha[1] * ((V(Xb).squeeze() - problem.g(Xb)) ** 2).mean()                # :669-670
    loss_boundary = loss.detach().clone()
    X = X0.clone().requires_grad_(True)
    Y = V(X).squeeze()                                                               # :713
    stopped = pt.zeros(K, dtype=pt.bool)
    V_L2 = pt.zeros(K, dtype=dtype)
    K_count = 0
    drawn = []
    for n in range(N):
        Y_ = V(X)
        grad_V, = pt.autograd.grad(Y_.squeeze().sum(), X, create_graph=True)         # :724
        Z = (problem.B.t() @ grad_V.t()).t()                                         # :725
        xin = xis[n] if xis is not None else pt.randn(K, problem.d).to(dtype)        # :726
        drawn.append(xin)
        sel = ~stopped
        if int(sel.sum()) == 0:
            break
        V_L2 = V_L2 + ((Y_.squeeze() - problem.v_true(X)) ** 2).detach() * dt * sel.to(dtype)   # :733
        X_prop = X + (problem.b(X) * dt + (problem.B @ xin.t()).t() * sq) * sel.to(dtype).unsqueeze(1)   # :741-742
        new_sel = elliptic_exit_mask(problem, X, X_prop)
        actf = (new_sel & ~stopped).to(dtype)
        Y = Y + (-problem.h(X, Y_.squeeze(), Z) * dt + (Z * xin).sum(1) * sq) * actf  # :768-769 (c = 0)
        X = X * (1 - actf).unsqueeze(1) + X_prop * actf.unsqueeze(1)                 # :772-773
        K_count += int(actf.sum())                                                   # :775-776
        stopped = stopped | (~new_sel & ~stopped)                                    # :778-779
    loss = loss + alpha[0] * ((V(X).squeeze() - Y) ** 2).mean()                      # :790
    loss.backward()
    grads = [q.grad.detach().clone() for q in params]
    for q in params:
        q.requires_grad_(False)
    return dict(loss=loss.detach(), loss_boundary=loss_boundary, grads=grads, K_count=K_count, X=X.detach(),
                Y=Y.detach(), V_L2=V_L2, stopped=stopped, xis=drawn)


def sample_sphere(K, d, radius, dtype=pt.float32):
    """Uniform sample on the sphere, solver.py:646-648."""
    X = pt.randn(K, d).to(dtype)
    return radius * X / pt.sqrt((X ** 2).sum(1)).unsqueeze(1)


def sample_square_boundary(K_boundary, d, X_l, X_r, one_boundary=False, dtype=pt.float32):
    """Boundary samples of the square, solver.py:655-665 (numpy shuffles first, then one pt.rand draw)."""
    h = int(K_boundary / 2)
    s = np.concatenate([np.ones(h)[:, np.newaxis], np.zeros([h, d - 1])], 1)
    np.apply_along_axis(np.random.shuffle, 1, s)
    a = np.concatenate([s, np.zeros([h, d])]).astype(bool)
    b = np.concatenate([np.zeros([h, d]), s]).astype(bool)
    Xb = ((X_r - X_l) * pt.rand(K_boundary, d) + X_l).to(dtype)
    Xb[pt.tensor(a)] = X_r if one_boundary else X_l
    Xb[pt.tensor(b)] = X_r
    return Xb


def elliptic_draws(problem, K, K_boundary, N):
    """The random draws of one EllipticSolver iteration in the reference's order (:646-665, :687-708, :726).
    N = 0 leaves the increments to elliptic_iteration (drawn step by step until every path has stopped)."""
    d = problem.d
    if problem.boundary == "sphere":
        Xb = sample_sphere(K_boundary, d, problem.boundary_distance)
        X0 = sample_ball(K, d, problem.boundary_distance)
    elif problem.boundary == "two_spheres":                                          # :650-654, :694-701: K shrinks
        Xb = pt.randn(K_boundary, d)
        radii = pt.tensor([problem.boundary_distance_1] * int(K_boundary / 2) + [problem.boundary_distance_2] * int(K_boundary / 2))
        Xb = radii.unsqueeze(1) * Xb / pt.sqrt((Xb ** 2).sum(1)).unsqueeze(1)
        X0 = sample_ball(K, d, problem.boundary_distance_2)
        X0 = X0[pt.sqrt((X0 ** 2).sum(1)) > problem.boundary_distance_1, :]
    else:
        Xb = sample_square_boundary(K_boundary, d, problem.X_l, problem.X_r, problem.one_boundary)
        X0 = (problem.X_r - problem.X_l) * pt.rand(K, d) + problem.X_l
    xis = pt.stack([pt.randn(X0.shape[0], d) for _ in range(N)]) if N > 0 else None
    return Xb, X0, xis


def elliptic_train_loop(problem, params, K, K_boundary, N, delta_t, L, lr, alpha=(1.0, 1.0), seed=42, times=None):
    """EllipticSolver.train (solver.py:628-809) with the reference's RNG use and per-module Adam."""
    import time
    pt.manual_seed(seed)
    np.random.seed(seed)
    opt = pt.optim.Adam(params, lr=lr)
    losses, kcounts = [], []
    for _ in range(L):
        t0 = time.time()
        Xb, X0, _ = elliptic_draws(problem, K, K_boundary, 0)
        o = elliptic_iteration(problem, params, Xb, X0, None, delta_t, N, alpha)
        for q, g in zip(params, o["grads"]):
            q.grad = g
        opt.step()
        losses.append(float(o["loss"]))
        kcounts.append(o["K_count"])
        if times is not None:
            times.append(time.time() - t0)
    return losses, kcounts


def sample_ball_uniform_square(K, d, radius, dtype=pt.float32):
    """GeneralSolver(uniform_square=True), solver.py:1041-1043: direction from the cube, radius U."""
    X = (pt.rand(K, d) * 2 - 1).to(dtype)
    return radius * X / pt.sqrt((X ** 2).sum(1)).unsqueeze(1) * pt.rand(K).to(dtype).unsqueeze(1)


def sample_ball(K, d, radius, dtype=pt.float32):
    """Uniform sample in the d-ball, draw order of solver.py:1045-1046."""
    X = pt.randn(K, d).to(dtype)
    return radius * X / pt.sqrt((X ** 2).sum(1)).unsqueeze(1) * (pt.rand(K).to(dtype).unsqueeze(1) ** (1 / d))


# --------------------------------------------------------------------------- importance sampling (next row f1)
def importance_sampling(problem, net, params, xis, delta_t, solver_delta_t, time_approx="inner", N_solver=None):
    """utilities.py:287-359 with control='approx'; xis (N,K,d) supplied by the caller.

    Returns (mean_IS, variance_IS, rel_error_IS).  The control net is addressed through Z_n(X, t)
    (solver.py:360-362: n = ceil(t / solver_delta_t))."""
    N, K, d = xis.shape
    dtype = xis.dtype
    sq = float(np.sqrt(delta_t))
    sdt = pt.as_tensor(solver_delta_t, dtype=dtype)
    X = problem.X_0.repeat(K, 1)
    ito = pt.zeros(K, dtype=dtype)
    rie = pt.zeros(K, dtype=dtype)
    fint = pt.zeros(K, dtype=dtype)
    with pt.no_grad():
        for n in range(N):
            xin = xis[n]
            nn_ = int(pt.ceil(pt.as_tensor(n * delta_t) / sdt))
            ut = -control_eval(net, params, X, nn_, sdt, time_approx, N_solver)
            X = X + (problem.b(X) + (problem.B @ ut.t()).t()) * delta_t + (problem.B @ xin.t()).t() * sq
            ito += (ut * xin).sum(1) * sq
            rie += (ut ** 2).sum(1) * delta_t
            fint += problem.f(X, n * delta_t) * delta_t
        w = pt.exp(-fint - problem.g(X)) * pt.exp(-ito - 0.5 * rie)
        mean = w.mean().item()
        var = pt.var(w).item()
    return mean, var, float(np.sqrt(var) / mean)


# --------------------------------------------------------------------------- whole training loops (CPU baseline)
def to_device(problem, params, device):
    """Move the tensors of a make_problem() namespace and a parameter list (or list of lists) to `device` in place: the
    reference's own device='cuda' mode (solver.py:36) -- same eager procedure, noise still drawn on the host (:381)."""
    for k, v in list(vars(problem).items()):
        if isinstance(v, pt.Tensor):
            setattr(problem, k, v.to(device))
    nets = params if isinstance(params[0], (list, tuple)) else [params]
    for net_ in nets:
        for i, q in enumerate(net_):
            net_[i] = q.detach().to(device)
    return problem, params


def hjb_train_loop(problem, net, params, K, delta_t, L, lr, loss_method="log-variance", time_approx="inner",
                   adaptive=True, detach_forward=True, seed=42, times=None, device=None):
    """Solver.train (solver.py:420-531) without logging extras: per iteration draw xi = randn(K, d, N+1)
    on the CPU generator (:381), roll out, loss, backward, one Adam step per network (:198-200)."""
    import time
    N = int(np.floor(problem.T / delta_t))                                          # :41
    nets = params if time_approx == "outer" else [params]
    for net_ in nets:
        for q in net_:
            q.requires_grad_(True)
    optims = [pt.optim.Adam(net_, lr=lr) for net_ in nets]
    pt.manual_seed(seed)                                                            # :422
    loss_log = []
    for _ in range(L):
        t0 = time.time()
        xi = pt.randn(K, problem.d, N + 1)
        if device is not None:
            xi = xi.to(device)                                                      # :381 `.to(self.device)`
        for o in optims:
            o.zero_grad()
        ad = True if loss_method == "relative_entropy" else adaptive
        X, Y, Zsum, _ = hjb_rollout(problem, net, params, xi, delta_t, N, time_approx, ad, detach_forward,
                                    want_zsum="relative_entropy" in loss_method)
        loss = hjb_loss(loss_method, problem, X, Y, Zsum, ad)
        loss.backward()
        for o in optims:
            o.step()
        loss_log.append(loss.item())
        if times is not None:
            if device is not None and pt.device(device).type == "cuda":
                pt.cuda.synchronize()
            times.append(time.time() - t0)
    return loss_log


def diffusion_train_loop(problem, params, K, K_boundary, N, delta_t, L, lr, seed=42, times=None):
    """GeneralSolver.train, loss 'diffusion', boundary 'unbounded' (solver.py:1001-1200)."""
    import time
    opt = pt.optim.Adam(params, lr=lr)
    pt.manual_seed(seed)                                                            # :1003
    loss_log, K_log = [], []
    for _ in range(L):
        t0 = time.time()
        X0 = sample_ball(K, problem.d, problem.boundary_distance)                   # :1045-1046
        t0_ = pt.rand(K, 1) * problem.T                                             # :1078
        # :1106 -- one draw per step, none once every path is stopped (break at :1093-1094; unbounded domain: the time test
        # of :1131 alone, replayed in fp32); the steps after the break get zeros, which diffusion_iteration never uses
        dt_, t_, stopped_, xl = pt.as_tensor(delta_t, dtype=pt.float32), t0_.squeeze(1).clone(), pt.zeros(K, dtype=pt.bool), []
        for _n in range(N):
            if int((~stopped_).sum()) == 0:
                break
            xl.append(pt.randn(K, problem.d))
            ns_ = (t_ + dt_) <= problem.T
            t_ = t_ + dt_ * (ns_ & ~stopped_).float()
            stopped_ = stopped_ | (~ns_ & ~stopped_)
        xis = pt.stack(xl + [pt.zeros(K, problem.d)] * (N - len(xl)))
        opt.zero_grad()
        out = diffusion_iteration(problem, params, X0, t0_, xis, delta_t, N, K_boundary)
        for q, g in zip(params, out["grads"]):
            q.grad = g
        opt.step()
        loss_log.append(float(out["loss"]))
        K_log.append(out["K_count"])
        if times is not None:
            times.append(time.time() - t0)
    return loss_log, K_log
