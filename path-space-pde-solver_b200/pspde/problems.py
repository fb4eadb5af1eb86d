"""Problem classes with the reference's interface (problems.py of lorenzrichter/path-space-PDE-solver) for the
problems on the fused hot path: LLGC (:14-65), LQGC (:118-175), DoubleWell (:178-282),
DoubleWell_multidim (:285-476), HeatEquation (:1733-1764).

Every class keeps the reference's constructor arguments, attributes (d, T, X_0, A, B, ...) and methods
(b, sigma, h, f, g, u_true, v_true) so that user code and the oracle comparisons read the same, and adds

    functor_pack() -> (problem_id, flags, fp32 vector)

which is how the problem reaches the CUDA kernels as a device functor (layout: include/pspde.h).  The torch
methods below exist for API compatibility and for host-side evaluation; the training hot path never calls them.
"""
import numpy as np
import torch as pt
from scipy.linalg import expm, solve_banded

from . import _lib as L


def default_device():
    return pt.device("cuda" if pt.cuda.is_available() else "cpu")


def _is_diagonal(M):
    M = M.detach().cpu()
    return bool(pt.equal(M, pt.diag(pt.diag(M))))


class _OrnsteinUhlenbeck:
    """Shared part of LLGC / LQGC: dX = (A X + B u) dt + B dW."""

    def _init_dynamics(self, d, off_diag, seed, device):
        self.device = default_device() if device is None else pt.device(device)
        pt.manual_seed(seed)                      # reference: problems.py:20 / :124 (global RNG side effect kept)
        self.A = (-pt.eye(d) + off_diag * pt.randn(d, d)).to(self.device)
        self.B = (pt.eye(d) + off_diag * pt.randn(d, d)).to(self.device)
        if not np.all(np.linalg.eigvals(self.A.cpu().numpy()).real < 0):
            print("not all EV of A are negative")

    def b(self, x):
        return x @ self.A.t()

    def sigma(self, x):
        return self.B

    def _pack(self, p_diag, r_diag, alpha):
        d = self.d
        z = pt.zeros(d)
        vecs = [pt.diag(self.A).cpu(), pt.diag(self.B).cpu(), p_diag, r_diag, alpha, z, z]
        flags = 0
        if not (_is_diagonal(self.A) and _is_diagonal(self.B)):
            flags |= L.FLAG_DENSE_AB
            vecs += [self.A.cpu().reshape(-1), self.B.cpu().reshape(-1)]
        return L.PROBLEM_OU, flags, pt.cat([v.float().reshape(-1) for v in vecs]).contiguous()


class LLGC(_OrnsteinUhlenbeck):
    """Ornstein-Uhlenbeck dynamics with linear terminal cost g(x) = alpha . x and no running state cost."""

    def __init__(self, name="LLGC", d=1, off_diag=0, T=5, seed=42, device=None):
        self.name, self.d, self.T = name, d, T
        self._init_dynamics(d, off_diag, seed, device)
        self.alpha = pt.ones(d, 1).to(self.device)
        self.X_0 = pt.zeros(d).to(self.device)
        self.boundary, self.one_boundary, self.X_l, self.X_r = "square", False, -2.0, 2.0

    def f(self, x, t):
        return pt.zeros(x.shape[0], device=x.device)

    def h(self, t, x, y, z):
        return -0.5 * (z * z).sum(dim=1)

    def g(self, x):
        return (x @ self.alpha)[:, 0]

    def u_true(self, x, t):
        """optimal control -B' exp(A'(T-t)) alpha, broadcast over the batch; returns (d, K) numpy like the reference."""
        A, B = self.A.cpu().numpy().astype(np.float64), self.B.cpu().numpy().astype(np.float64)
        col = -B.T @ (expm(A.T * (self.T - t)) @ self.alpha.cpu().numpy().astype(np.float64))
        return col * np.ones((self.d, x.shape[0]))

    def v_true(self, x, t):
        """value function; the covariance integral uses the reference's left Riemann sum (step 1e-3, :56-63)."""
        A, B = self.A.cpu().numpy().astype(np.float64), self.B.cpu().numpy().astype(np.float64)
        al = self.alpha.cpu().numpy().astype(np.float64)
        step = 0.001
        n = int(np.floor((self.T - t) / step)) + 1
        S = np.zeros((self.d, self.d))
        for s in np.linspace(t, self.T, n):
            E = expm(A * (self.T - s))
            S += E @ B @ B.T @ E.T * step
        xn = x.detach().cpu().numpy().astype(np.float64)
        return (expm(A * (self.T - t)) @ xn.T).T @ al - 0.5 * al.T @ S @ al

    def functor_pack(self):
        z = pt.zeros(self.d)
        return self._pack(z, z, self.alpha.cpu().reshape(-1))

    def u_true_table(self, N, delta_t):
        """Device-table form of u_true on the solver grid t_n = n * delta_t (include/pspde.h, pspde_udiag mode 1):
        u*(x, t_n) = U0[n] (independent of x).  Returns dict(mode, table (N, 2, d))."""
        A, B = self.A.cpu().numpy().astype(np.float64), self.B.cpu().numpy().astype(np.float64)
        al = self.alpha.cpu().numpy().astype(np.float64)
        tab = np.zeros((N, 2, self.d), np.float32)
        for n in range(N):
            tab[n, 0] = (-B.T @ (expm(A.T * (self.T - n * delta_t)) @ al))[:, 0]
        return dict(mode=1, table=pt.from_numpy(tab))


class LQGC(_OrnsteinUhlenbeck):
    """Linear-quadratic Gaussian control: running cost x'Px, terminal cost x'Rx (P = Q = I/2, R = I)."""

    def __init__(self, name="LQGC", delta_t=0.05, d=1, off_diag=0, T=5, seed=42, device=None):
        self.name, self.d, self.T = name, d, T
        self._init_dynamics(d, off_diag, seed, device)
        self.delta_t = delta_t
        self.N = int(np.floor(self.T / self.delta_t))
        self.X_0 = pt.zeros(d).to(self.device)
        eye = pt.eye(d).to(self.device)
        self.P, self.Q, self.R = 0.5 * eye, 0.5 * eye, eye.clone()
        # backward Riccati recursion for V(x, t_n) = -x'F_n x + G_n on the problem's own grid (:140-152)
        Qi_Bt = pt.linalg.solve(self.Q, self.B.t())
        F = pt.zeros(self.N + 1, d, d, device=self.device)
        G = pt.zeros(self.N + 1)
        F[self.N] = self.R
        for n in range(self.N, 0, -1):
            Fn = F[n]
            F[n - 1] = Fn + (self.A.t() @ Fn + Fn @ self.A - Fn @ self.B @ Qi_Bt @ Fn + self.P) * delta_t
            G[n - 1] = G[n] - pt.trace(self.B @ Fn @ self.B).cpu() * delta_t
        self.F, self.G = F, G

    def f(self, x, t):
        return ((x @ self.P.t()) * x).sum(dim=1)

    def g(self, x):
        return ((x @ self.R.t()) * x).sum(dim=1)

    def h(self, t, x, y, z):
        return -0.5 * (z * z).sum(dim=1) - self.f(x, t)

    def u_true(self, x, t):
        n = int(np.ceil(t / self.delta_t))
        gain = pt.linalg.solve(self.Q, self.B.t()).cpu() @ self.F[n].cpu()
        return -(gain @ x.detach().cpu().t()).numpy()

    def v_true(self, x, t):
        n = int(np.ceil(t / self.delta_t))
        return -(x @ (self.F[n] @ x.t())).t() + self.G[n]

    def functor_pack(self):
        for M, nm in ((self.P, "P"), (self.R, "R")):
            if not _is_diagonal(M):
                raise NotImplementedError("LQGC with non-diagonal %s is not supported by the fused kernels" % nm)
        return self._pack(pt.diag(self.P).cpu(), pt.diag(self.R).cpu(), pt.zeros(self.d))

    def u_true_table(self, N, delta_t):
        """u*(x, t_n) = -Q^-1 B' F[ceil(t_n / self.delta_t)] x; table form needs a diagonal gain (else None)."""
        tab = np.zeros((N, 2, self.d), np.float32)
        Qi_Bt = pt.linalg.solve(self.Q, self.B.t()).cpu()
        for n in range(N):
            k = int(np.ceil(n * delta_t / self.delta_t))
            gain = Qi_Bt @ self.F[min(k, self.N)].cpu()
            if not pt.equal(gain, pt.diag(pt.diag(gain))):
                return None
            tab[n, 1] = -pt.diag(gain).numpy()
        return dict(mode=1, table=pt.from_numpy(tab))


# ----------------------------------------------------------------------------------------------- double well
def _fd_reference(V, g_term, T, delta_t, xb, nx, B00=1.0):
    """Finite-difference reference for psi = exp(-V_value) of the 1-d double-well HJB (implicit Euler in time,
    symmetrised generator with Neumann ends; the scheme of problems.py:216-268, vectorised).

    Returns (psi (N+1, nx), u (N+1, nx-1), dx)."""
    beta = 2.0
    dx = 2.0 * xb / nx
    i = np.arange(nx)
    xm = -xb + (i + 0.5) * dx                      # cell mid points
    xl, xr = -xb + i * dx, -xb + (i + 1) * dx      # cell faces
    lo = np.exp(beta * 0.5 * (V(xm - dx) + V(xm) - 2 * V(xl))) / dx ** 2
    up = np.exp(beta * 0.5 * (V(xm + dx) + V(xm) - 2 * V(xr))) / dx ** 2
    dl = np.exp(beta * (V(xm) - V(xl))) / dx ** 2
    du = np.exp(beta * (V(xm) - V(xr))) / dx ** 2
    diag = np.where(i > 0, dl, 0.0) + np.where(i < nx - 1, du, 0.0)
    A_diag = -diag / beta
    A_up = up[:-1] / beta                          # A[i, i+1]
    N = int(T / delta_t)
    xvec = np.linspace(-xb, xb, nx, endpoint=True)
    Dh = np.exp(beta * V(xvec) / 2)
    band = -delta_t * np.vstack([np.append([0.0], A_up), A_diag - N / T, np.append(A_up, [0.0])])
    psi = np.zeros((N + 1, nx))
    psi[N] = np.exp(-g_term(xvec))
    for n in range(N - 1, -1, -1):
        psi[n] = Dh * solve_banded((1, 1), band, psi[n + 1] / Dh)
    lp = np.log(psi)
    u = -2.0 / beta * B00 * (lp[:, :-1] - lp[:, 1:]) / dx
    return psi, u, dx, xvec


def _grid_index(x, xb, dx, clip=True):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if clip:
        x = np.clip(x, -xb, xb - 2 * dx)
    idx = np.floor((x + xb) / dx).astype(np.int64)
    idx[-1] -= 2                                   # reference quirk (:273, :279), kept for identical outputs
    return idx


class DoubleWell:
    """One-dimensional double well V(x) = kappa (x^2 - 1)^2, terminal cost eta (x - 1)^2."""

    def __init__(self, name="Double well", d=1, T=1, eta=1, kappa=1, device=None):
        self.device = default_device() if device is None else pt.device(device)
        self.name, self.d, self.T, self.eta, self.kappa = name, d, T, eta, kappa
        self.B = pt.eye(d).to(self.device)
        self.X_0 = -pt.ones(d).to(self.device)
        self.ref_sol_is_defined = False
        if d != 1:
            print("The double well example is only implemented for d = 1.")

    def V(self, x):
        return self.kappa * (x ** 2 - 1) ** 2

    def grad_V(self, x):
        return 4.0 * self.kappa * x * (x ** 2 - 1)

    def b(self, x):
        return -self.grad_V(x)

    def sigma(self, x):
        return self.B

    def f(self, x, t):
        return pt.zeros(x.shape[0], device=x.device)

    def h(self, t, x, y, z):
        return -0.5 * (z * z).sum(dim=1)

    def g(self, x):
        return (self.eta * (x - 1) ** 2).squeeze()

    def compute_reference_solution(self, delta_t=0.005, xb=2.5, nx=1000):
        self.xb, self.nx, self.delta_t = xb, nx, delta_t
        self.psi, self.u, self.dx, self.xvec = _fd_reference(self.V, lambda x: self.eta * (x - 1) ** 2, self.T,
                                                             delta_t, xb, nx, float(self.B[0, 0]))
        self.ref_sol_is_defined = True

    def v_true(self, x, t):
        idx = _grid_index(x, self.xb, self.dx, clip=False)
        n = int(np.ceil(t / self.delta_t))
        return -np.log(self.psi[n, idx]).reshape(1, -1)

    def u_true(self, x, t):
        idx = _grid_index(x, self.xb, self.dx)
        n = int(np.ceil(t / self.delta_t))
        return self.u[n, idx].reshape(1, -1)

    def functor_pack(self):
        d = self.d
        z, one = pt.zeros(d), pt.ones(d)
        return L.PROBLEM_DW, 0, pt.cat([z, one, z, z, z, self.kappa * one, self.eta * one]).float().contiguous()

    def u_true_table(self, N, delta_t):
        """pspde_udiag mode 2 from the finite-difference reference (compute_reference_solution must have run).
        quirk_last: the reference shifts the table cell of the LAST batch element by -2 (`i[-1] -= 2`, :279); the kernels
        reproduce it for the path with global index K - 1 (pspde_udiag.quirk_path)."""
        if not hasattr(self, "u"):
            return None
        rows = [min(int(np.ceil(n * delta_t / self.delta_t)), self.u.shape[0] - 1) for n in range(N)]
        tab = np.stack([np.stack([self.u[r], self.u[r]]) for r in rows]).astype(np.float32)
        return dict(mode=2, table=pt.from_numpy(tab), nx1=self.u.shape[1], d1=self.d, xb=self.xb, dx=self.dx, quirk_last=True)


class DoubleWell_multidim:
    """d independent double wells; the first d_1 coordinates use (kappa, eta), the remaining d_2 use (1, 1)."""

    def __init__(self, name="Double well", d=1, d_1=1, d_2=0, T=1, eta=1, kappa=1, device=None):
        self.device = default_device() if device is None else pt.device(device)
        self.name, self.d, self.d_1, self.d_2, self.T = name, d, d_1, d_2, T
        self.eta, self.kappa = eta, kappa
        self.eta_ = pt.tensor([eta] * d_1 + [1.0] * d_2).to(self.device)
        self.kappa_ = pt.tensor([kappa] * d_1 + [1.0] * d_2).to(self.device)
        self.B = pt.eye(d).to(self.device)
        self.X_0 = -pt.ones(d).to(self.device)
        self.ref_sol_is_defined = False
        self.boundary, self.boundary_distance = "unbounded", 2.0

    def V(self, x):
        return self.kappa * (x ** 2 - 1) ** 2

    def V_2(self, x):
        return (x ** 2 - 1) ** 2

    def grad_V(self, x):
        return 4.0 * self.kappa_ * (x * (x ** 2 - 1.0))

    def b(self, x):
        return -self.grad_V(x)

    def sigma(self, x):
        return self.B

    def h(self, t, x, y, z):
        return -0.5 * (z * z).sum(dim=1)

    def f(self, x, t):
        return pt.zeros(x.shape[0], device=x.device)

    def g_1(self, x_1):
        return self.eta * (x_1 - 1) ** 2

    def g_2(self, x_1):
        return (x_1 - 1) ** 2

    def g(self, x):
        return (self.eta_ * (x - 1.0) ** 2).sum(dim=1).squeeze()

    def compute_reference_solution(self, delta_t=0.005, xb=2.5, nx=1000):
        self.xb, self.nx, self.delta_t = xb, nx, delta_t
        self.psi, self.u, self.dx, self.xvec = _fd_reference(self.V, self.g_1, self.T, delta_t, xb, nx,
                                                             float(self.B[0, 0]))

    def compute_reference_solution_2(self, delta_t=0.005, xb=2.5, nx=1000):
        self.xb, self.nx, self.delta_t = xb, nx, delta_t
        self.psi_2, self.u_2, self.dx, self.xvec = _fd_reference(self.V_2, self.g_2, self.T, delta_t, xb, nx,
                                                                 float(self.B[0, 0]))

    def v_true_1(self, x, t):
        idx = _grid_index(x, self.xb, self.dx, clip=False)
        return -np.log(self.psi[int(np.ceil(t / self.delta_t)), idx]).reshape(1, -1)

    def u_true_1(self, x, t):
        idx = _grid_index(x, self.xb, self.dx)
        return self.u[int(np.ceil(t / self.delta_t)), idx].reshape(1, -1)

    def u_true_2(self, x, t):
        idx = _grid_index(x, self.xb, self.dx)
        return self.u_2[int(np.ceil(t / self.delta_t)), idx].reshape(1, -1)

    def v_true(self, x, t):
        return None

    def u_true(self, x, t):
        x = x.detach().cpu() if isinstance(x, pt.Tensor) else x
        cols = [self.u_true_1(x[:, i], t).T for i in range(self.d_1)]
        cols += [self.u_true_2(x[:, i], t).T for i in range(self.d_1, self.d)]
        return np.concatenate(cols, 1).T

    def functor_pack(self):
        d = self.d
        z, one = pt.zeros(d), pt.ones(d)
        return L.PROBLEM_DW, 0, pt.cat([z, one, z, z, z, self.kappa_.cpu(), self.eta_.cpu()]).float().contiguous()

    def u_true_table(self, N, delta_t):
        """pspde_udiag mode 2: class 0 = the first d_1 coordinates (u), class 1 = the rest (u_2)."""
        if not hasattr(self, "u") or (self.d_2 > 0 and not hasattr(self, "u_2")):
            return None
        u2 = self.u_2 if self.d_2 > 0 else self.u
        rows = [min(int(np.ceil(n * delta_t / self.delta_t)), self.u.shape[0] - 1) for n in range(N)]
        tab = np.stack([np.stack([self.u[r], u2[r]]) for r in rows]).astype(np.float32)
        return dict(mode=2, table=pt.from_numpy(tab), nx1=self.u.shape[1], d1=self.d_1, xb=self.xb, dx=self.dx,
                    quirk_last=True)                       # `i[-1] -= 2` in u_true_1 / u_true_2 (:401, :464)


class HeatEquation:
    """Backward heat equation d_t V + Laplace V = 0, V(x, T) = |x|^2; exact V = |x|^2 + 2 d (T - t)."""

    def __init__(self, name="Heat equation", d=1, T=1, seed=42, device=None):
        self.device = default_device() if device is None else pt.device(device)
        pt.manual_seed(seed)
        self.name, self.d, self.T = name, d, T
        self.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(self.device)
        self.boundary, self.boundary_type, self.boundary_distance = "unbounded", "Dirichlet", 1.0

    def b(self, x):
        return pt.zeros_like(x)

    def sigma(self, x):
        return self.B

    def g(self, x, t):
        return pt.zeros(x.shape[0], device=x.device)

    def h(self, t, x, y, z):
        return pt.zeros(x.shape[0], device=x.device)

    def f(self, x):
        return (x * x).sum(dim=1)

    def u_true(self, x, t):
        return None

    def v_true(self, x, t):
        return (x * x).sum(dim=1) + 2 * (self.T - t) * self.d

    def functor_pack(self):
        d = self.d
        z = pt.zeros(d)
        return L.PROBLEM_HEAT, 0, pt.cat([z, pt.diag(self.B).cpu(), z, z, z, z, z]).float().contiguous()


class AllenCahn:
    """Allen-Cahn equation d_t V + Laplace V + V - V^3 = 0, V(x, T) = 1 / (2 + 2/5 |x|^2) (problems.py:1175-1218; the
    d = 100 benchmark of the 'Allen-Cahn' notebook).  The reference's numpy / torch switch ``modus`` is kept as an
    attribute for notebook compatibility; everything here is torch on the device."""

    def __init__(self, name="Allen-Cahn", d=1, T=0.3, seed=42, modus="pt", device=None):
        self.device = default_device() if device is None else pt.device(device)
        np.random.seed(seed)
        self.modus, self.name, self.d, self.T = modus, name, d, T
        self.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(self.device)
        self.B_pt = self.B
        self.X_0 = np.zeros(d)
        self.sigma_modus, self.boundary, self.boundary_distance = "constant", "unbounded", 2.0

    def b(self, x):
        return pt.zeros_like(x)

    def sigma(self, x):
        return self.B

    def h(self, t, x, y, z):
        return y - y ** 3

    def f(self, x):
        return 1 / (2 + 2 / 5 * (x ** 2).sum(1))

    def functor_pack(self):
        z = pt.zeros(self.d)
        return L.PROBLEM_ALLEN_CAHN, 0, pt.cat([z, pt.diag(self.B).cpu(), z, z, z, z, z]).float().contiguous()


# ---------------------------------------------------------------------------------------------- elliptic problems
class _ExponentialOnBall:
    """Common part of the elliptic toy problems on the unit ball (problems.py:962-1064): b = 0, sigma = sqrt(2) I,
    Dirichlet data and exact solution exp(alpha |x|^2).  Subclasses supply h(x, y, z) and its kernel functor id."""
    H_ID = L.H_ZERO

    def __init__(self, name, d=2, alpha=1.0, boundary_type="Dirichlet", device=None):
        self.device = default_device() if device is None else pt.device(device)
        self.name, self.d, self.alpha = name, d, alpha
        self.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(self.device)
        self.X_0 = pt.zeros(d, device=self.device)
        self.Y_0 = pt.zeros(1, device=self.device)
        self.boundary, self.boundary_distance, self.boundary_type = "sphere", 1.0, boundary_type

    def b(self, x):
        return pt.zeros_like(x)

    def sigma(self, x):
        return self.B

    def f(self, x, t=None):
        return pt.zeros(x.shape[0], device=x.device)

    def g(self, x):
        if self.boundary_type == "Neumann":
            return 2 * self.alpha * x * pt.exp(self.alpha * (x ** 2).sum(1)).unsqueeze(1)
        return pt.exp(self.alpha * (x ** 2).sum(1))

    def u_true(self, x):
        return -2 * np.sqrt(2.0) * self.alpha * x * pt.exp(self.alpha * (x ** 2).sum(1).unsqueeze(1))

    def v_true(self, x):
        return pt.exp(self.alpha * (x ** 2).sum(1))

    def functor_pack(self):
        z = pt.zeros(self.d)
        return L.PROBLEM_HEAT, 0, pt.cat([z, pt.diag(self.B).cpu(), z, z, z, z, z]).float().contiguous()

    def elliptic_spec(self):
        """Domain + h functor for the kernels (include/pspde.h: pspde_elliptic)."""
        return L.make_elliptic(L.DOMAIN_SPHERE, radius=self.boundary_distance, h_id=self.H_ID,
                               h_param=(self.alpha, 0.0, 0.0))


class ExponentialOnSphere(_ExponentialOnBall):
    """problems.py:962-992: linear h = -alpha y (4 alpha |x|^2 + 2 d)."""
    H_ID = L.H_EXP_LINEAR

    def __init__(self, name="Exponential on sphere", d=2, alpha=1.0, device=None):
        super().__init__(name, d, alpha, "Dirichlet", device)

    def h(self, x, y, z):
        return -self.alpha * y * (self.alpha * 4 * (x ** 2).sum(1) + 2 * self.d)


class ExponentialOnBallNonlinear(_ExponentialOnBall):
    """problems.py:995-1028: h = -2 alpha y (2 alpha |x|^2 + d) + exp(2 alpha |x|^2) - y^2."""
    H_ID = L.H_EXP_NONLINEAR

    def __init__(self, name="Exponential on ball nonlinear", d=2, alpha=1.0, boundary_type="Dirichlet", device=None):
        super().__init__(name, d, alpha, boundary_type, device)

    def h(self, x, y, z):
        r2 = (x ** 2).sum(1)
        return -2 * self.alpha * y * (self.alpha * 2 * r2 + self.d) + pt.exp(2 * self.alpha * r2) - y ** 2


class ExponentialOnBallNonlinearSin(_ExponentialOnBall):
    """problems.py:1031-1064: h = -2 alpha y (2 alpha |x|^2 + d) + sin(exp(2 alpha |x|^2) - y^2)."""
    H_ID = L.H_EXP_NONLINEAR_SIN

    def __init__(self, name="Exponential on ball nonlinear", d=2, alpha=1.0, boundary_type="Dirichlet", device=None):
        super().__init__(name, d, alpha, boundary_type, device)

    def h(self, x, y, z):
        r2 = (x ** 2).sum(1)
        return -2 * self.alpha * y * (self.alpha * 2 * r2 + self.d) + pt.sin(pt.exp(2 * self.alpha * r2) - y ** 2)


class Helmholtz:
    """problems.py:1614-1654: Helmholtz equation on the square [-1, 1]^2 with exact solution
    sin(a_1 pi x_0) sin(a_2 pi x_1)."""

    def __init__(self, name="Helmholtz", d=2, r=1.0, device=None):
        self.device = default_device() if device is None else pt.device(device)
        self.name, self.d = name, d
        self.B = (pt.sqrt(pt.tensor(2.0)) * pt.eye(d)).to(self.device)
        self.X_0 = -pt.ones(d, device=self.device)
        self.a_1, self.a_2, self.k = 1.0, 4.0, 1.0
        self.pi = pt.tensor(np.pi)
        self.boundary, self.one_boundary, self.X_l, self.X_r = "square", False, -1.0, 1.0
        if d != 2:
            print("Only implemented for d = 2.")

    def _ss(self, x):
        return pt.sin(self.a_1 * self.pi * x[:, 0]) * pt.sin(self.a_2 * self.pi * x[:, 1])

    def b(self, x):
        return pt.zeros_like(x)

    def sigma(self, x):
        return self.B

    def f(self, x):
        return pt.zeros(x.shape[0], device=x.device)

    def g(self, x):
        return self._ss(x)

    def h(self, x, y, z):
        s = self._ss(x)
        return self.k ** 2 * y + (self.a_1 * self.pi) ** 2 * s + (self.a_2 * self.pi) ** 2 * s - self.k ** 2 * s

    def v_true(self, x):
        return self._ss(x)

    def functor_pack(self):
        z = pt.zeros(self.d)
        return L.PROBLEM_HEAT, 0, pt.cat([z, pt.diag(self.B).cpu(), z, z, z, z, z]).float().contiguous()

    def elliptic_spec(self):
        return L.make_elliptic(L.DOMAIN_BOX, x_l=self.X_l, x_r=self.X_r, one_boundary=self.one_boundary,
                               h_id=L.H_HELMHOLTZ, h_param=(self.k, self.a_1, self.a_2))


class Committor:
    """problems.py:1546-1580: committor function between two concentric spheres a < |x| < c (Brownian motion, sigma = I,
    h = 0, boundary data 0 on the inner and 1 on the outer sphere); exact solution harmonic in r."""

    def __init__(self, name="Committor", d=2, alpha=1.0, device=None):
        self.device = default_device() if device is None else pt.device(device)
        self.name, self.d = name, d
        self.a, self.c = 1.0, 2.0
        self.B = pt.eye(d).to(self.device)
        self.X_0 = pt.zeros(d, device=self.device)
        self.Y_0 = pt.zeros(1, device=self.device)
        self.boundary, self.boundary_distance_1, self.boundary_distance_2 = "two_spheres", self.a, self.c

    def b(self, x):
        return pt.zeros_like(x)

    def sigma(self, x):
        return self.B

    def f(self, x):
        return pt.zeros(x.shape[0], device=x.device)

    def g(self, x):
        return (pt.sqrt((x ** 2).sum(1)) > self.a).float()

    def h(self, x, y, z):
        return pt.zeros(x.shape[0], device=x.device)

    def u_true(self, x):
        return pt.zeros(x.shape)

    def v_true(self, x):
        return ((self.a ** 2 - pt.sqrt((x ** 2).sum(1)) ** (2 - self.d) * self.a ** self.d)
                / (self.a ** 2 - self.c ** (2 - self.d) * self.a ** self.d))

    def functor_pack(self):
        z = pt.zeros(self.d)
        return L.PROBLEM_HEAT, 0, pt.cat([z, pt.diag(self.B).cpu(), z, z, z, z, z]).float().contiguous()

    def elliptic_spec(self):
        return L.make_elliptic(L.DOMAIN_ANNULUS, radius=self.boundary_distance_2, radius_in=self.boundary_distance_1,
                               h_id=L.H_COMMITTOR, h_param=(self.a, self.c, 0.0))
