"""EllipticSolver with the interface of the reference's ``EllipticSolver`` (solver.py:560-931) for the diffusion loss
on a bounded domain: random interior start points, diffusion stopped at the sphere / square
boundary, Dirichlet boundary term.  The N-step rollout with the directional derivative of V, the exit masks, the
h(x, V(x)) term, the V_L2 diagnostic, the loss and its gradient run in the fused sm_100a kernels of
csrc/diffusion_kernels.cuh behind pspde_elliptic_* (include/pspde.h).

Kept from the reference: the constructor signature (:562-566), the attribute ``V`` (a DenseNet on X only, :606, which
the caller may replace before ``train()``), ``train()`` (:628), the per-module Adam (``V.optim``, :797) and the
result lists ``loss_log, K_log, V_L2_log, V_test_L2, V_test_abs, V_test_rel_abs, times`` (:612-626).
Added keyword arguments as in ``pspde.GeneralSolver``: ``noise='philox' | 'inject'`` (inject = the reference's CPU
draw order: boundary samples :646-665, start points :687-708, one randn(K, d) per step :726), ``device``,
``process_group`` (K is the GLOBAL batch, sharded over ranks).
Options off this path raise NotImplementedError (other losses incl. PINN / BSDE, approx_method='Z', Neumann
boundary term, 'two_spheres' / 'square-corner' domains, adaptive or attached forward process, sample_center,
loss_with_stopped, variance_moment_split, uniform_square).  There is no CPU fallback.
"""
import ctypes
import time
from datetime import date

import numpy as np
import torch as pt

from . import _lib as L
from . import dist
from .function_space import DenseNet
from .fused import on_own_device
from .general_solver import DiffusionCall, DiffusionEngine, FusedDiffusion, GeneralSolver


class EllipticEngine(DiffusionEngine):
    """Device buffers + the library calls of the elliptic rollout for one (problem, network, K_local, N)."""

    def __init__(self, problem, dims, K_local, N, delta_t, k_offset=0, seed=42, device=None):
        self.ell = problem.elliptic_spec()
        super().__init__(problem, dims, K_local, N, delta_t, k_offset=k_offset, seed=seed, device=device)
        self.VL2 = pt.zeros(self.K_local, dtype=pt.float32, device=self.device)

    def cfg(self, K, N, xis, offset):
        noise, strides = L.NOISE_PHILOX, (0, 0, 0)
        if xis is not None:
            if xis.dtype != pt.float32 or xis.device != self.device or tuple(xis.shape) != (N, K, self.d):
                raise ValueError("xis must be a float32 (N, K_local, d) tensor on %s" % self.device)
            noise, strides = L.NOISE_INJECT, (xis.stride(1), xis.stride(2), xis.stride(0))
        return L.make_cfg(K, self.d, N, self.dt, self.problem_id, L.NET_DENSENET, self.dims, L.TIME_NONE,
                          adaptive=False, k_offset=self.k_offset, problem_flags=self.flags, noise_mode=noise,
                          seed=self.seed, offset=offset, xi_strides=strides, n_sets=1)

    def workspace_bytes(self, cfg):
        return int(self.lib.pspde_elliptic_workspace_bytes(ctypes.byref(cfg), ctypes.byref(self.ell)))

    @on_own_device
    def forward(self, theta, X0, t0, xis, offset, N=None, outs=None):
        K = X0.shape[0]
        N = self.N if N is None else N
        cfg = self.cfg(K, N, xis if N > 0 else None, offset)
        full = outs is None
        V0, VE, Y, X_end, _, stats = (self.V0, self.VE, self.Y, self.X_end, None, self.stats) if full else outs
        VL2 = self.VL2 if full else None
        if not full and N > 0:                       # a batch of another size ('two_spheres': K changes every iteration)
            self.VL2_last = VL2 = pt.zeros(K, dtype=pt.float32, device=self.device)
            self.stats_last = stats = pt.zeros(4, dtype=pt.float64, device=self.device)
        elif full:
            self.VL2_last, self.stats_last = self.VL2, self.stats
        rc = self.lib.pspde_elliptic_fwd(ctypes.byref(cfg), ctypes.byref(self.ell), self._p(theta), self._p(self.pack),
                                         self._p(X0), self._p(xis if N > 0 else None), self._p(V0), self._p(VE),
                                         self._p(Y), self._p(X_end), self._p(VL2),
                                         self._p(stats), self._p(self.workspace), self.workspace.numel(), self._stream())
        L.check(self.lib, rc)

    @on_own_device
    def backward(self, theta, X0, t0, xis, offset, c0, cE, cD, grad_out, N=None):
        K = X0.shape[0]
        N = self.N if N is None else N
        cfg = self.cfg(K, N, xis if N > 0 else None, offset)
        rc = self.lib.pspde_elliptic_bwd(ctypes.byref(cfg), ctypes.byref(self.ell), self._p(theta), self._p(self.pack),
                                         self._p(X0), self._p(xis if N > 0 else None), self._p(c0), self._p(cE),
                                         self._p(cD), self._p(grad_out), self._p(self.workspace),
                                         self.workspace.numel(), self._stream())
        L.check(self.lib, rc)


class EllipticSolver(GeneralSolver):

    def __init__(self, problem, name, seed=42, delta_t=0.01, N=50, lr=0.001, L=100000, K=200, K_boundary=50,
                 alpha=[1.0, 1.0], adaptive_forward_process=False, detach_forward=True, print_every=100, verbose=True,
                 approx_method='Y', sample_center=False, loss_method='diffusion', loss_with_stopped=False,
                 K_test_log=None, PINN_log_variance=False, log_loss_parts=False, boundary_loss=True,
                 boundary_type='Dirichlet', variance_moment_split=False, full_hessian=False, uniform_square=False,
                 noise='philox', device=None, process_group=None):
        self.problem, self.name = problem, name
        self.date = date.today().strftime('%Y-%m-%d')
        self.d = problem.d
        self.device = pt.device('cuda', pt.cuda.current_device()) if device is None else pt.device(device)
        self.seed = seed
        self.delta_t_np = delta_t
        self.delta_t = pt.tensor(self.delta_t_np).to(self.device)
        self.sq_delta_t = pt.sqrt(self.delta_t).to(self.device)
        self.N, self.lr, self.L, self.K, self.K_original, self.K_boundary = N, lr, L, K, K, K_boundary
        self.alpha = list(alpha)
        self.boundary_type = boundary_type
        self.adaptive_forward_process, self.detach_forward = adaptive_forward_process, detach_forward
        self.approx_method, self.sample_center, self.loss_method = approx_method, sample_center, loss_method
        self.loss_with_stopped, self.boundary_loss = loss_with_stopped, boundary_loss
        self.PINN_log_variance, self.variance_moment_split = PINN_log_variance, variance_moment_split
        self.full_hessian, self.uniform_square = full_hessian, uniform_square
        self.print_every, self.verbose = print_every, verbose
        self.noise, self.process_group = noise, process_group
        off = []
        if loss_method != 'diffusion':
            off.append('loss_method=%r' % loss_method)
        if approx_method != 'Y':
            off.append('approx_method=%r' % approx_method)
        if adaptive_forward_process or not detach_forward:
            off.append('adaptive / attached forward process')
        if getattr(problem, 'boundary', None) not in ('sphere', 'square', 'two_spheres'):
            off.append('boundary=%r' % getattr(problem, 'boundary', None))
        if getattr(problem, 'boundary', None) == 'two_spheres' and process_group is not None:
            off.append("'two_spheres' (batch size changes every iteration) with a process group")
        if boundary_loss and boundary_type != 'Dirichlet':
            off.append('boundary_type=%r' % boundary_type)
        if sample_center or loss_with_stopped or uniform_square or variance_moment_split or full_hessian \
                or PINN_log_variance:
            off.append('sample_center / loss_with_stopped / uniform_square / variance_moment_split / full_hessian')
        if not hasattr(problem, 'elliptic_spec'):
            off.append('problem %s has no device functor (elliptic_spec)' % type(problem).__name__)
        if off:
            raise NotImplementedError("off the fused hot path: " + ", ".join(off))
        if noise not in ('philox', 'inject'):
            raise ValueError("noise must be 'philox' or 'inject'")
        pt.manual_seed(seed)                                      # solver.py:604
        self.V = DenseNet(d_in=self.d, d_out=1, lr=self.lr, seed=seed).to(self.device)     # :606
        self.K_test_log = K_test_log
        self.Y_0_log, self.loss_log, self.loss_log_domain, self.loss_log_boundary = [], [], [], []
        self.u_L2_log, self.V_L2_log, self.V_test_L2, self.V_test_abs, self.V_test_rel_abs = [], [], [], [], []
        self.times, self.lambda_log, self.K_log, self.path_steps_per_sec = [], [], [], []
        self.log_loss_parts = log_loss_parts
        self._engine, self._V_homed, self._iteration = None, None, 0

    def _get_engine(self):
        self._home_parameters()
        if self._engine is None:
            net_id, dims = self.V.net_spec()
            if net_id != L.NET_DENSENET or dims[0] != self.d or dims[-1] != 1:
                raise NotImplementedError("the elliptic kernels need V = DenseNet(d_in=d, d_out=1)")
            rank, W = dist.world(self.process_group)
            self._k_lo, self._k_hi = dist.shard_range(self.K, rank, W)
            self._engine = EllipticEngine(self.problem, dims, self._k_hi - self._k_lo, self.N, self.delta_t_np,
                                          k_offset=self._k_lo, seed=self.seed, device=self.device)
            if self._engine.n_theta != self._theta.numel():
                raise RuntimeError("parameter count mismatch: module %d vs kernel %d"
                                   % (self._theta.numel(), self._engine.n_theta))
        return self._engine

    # ------------------------------------------------------------------ sampling (solver.py:646-665, :687-708)
    def _sample_boundary_cpu(self):
        p, Kb, d = self.problem, self.K_boundary, self.d
        if p.boundary == 'sphere':
            Xb = pt.randn(Kb, d)
            return p.boundary_distance * Xb / pt.sqrt(pt.sum(Xb ** 2, 1)).unsqueeze(1)
        if p.boundary == 'two_spheres':                        # solver.py:650-654: half of the samples on each sphere
            Xb = pt.randn(Kb, d)
            radii = pt.tensor([p.boundary_distance_1] * int(Kb / 2) + [p.boundary_distance_2] * int(Kb / 2))
            return radii.unsqueeze(1) * Xb / pt.sqrt(pt.sum(Xb ** 2, 1)).unsqueeze(1)
        h = int(Kb / 2)                                        # 'square': numpy shuffles, then one pt.rand draw
        s = np.concatenate([np.ones(h)[:, np.newaxis], np.zeros([h, d - 1])], 1)
        np.apply_along_axis(np.random.shuffle, 1, s)
        a = np.concatenate([s, np.zeros([h, d])]).astype(bool)
        b = np.concatenate([np.zeros([h, d]), s]).astype(bool)
        Xb = (p.X_r - p.X_l) * pt.rand(Kb, d) + p.X_l
        Xb[pt.tensor(a)] = p.X_r if p.one_boundary else p.X_l
        Xb[pt.tensor(b)] = p.X_r
        return Xb

    def _draw_increments_cpu(self, X):
        """'inject' only: one randn(K, d) per step, drawn BEFORE the all-stopped check and not at all afterwards
        (solver.py:726-730), so that a whole training loop consumes the CPU RNG stream exactly like the reference.
        The exit times do not depend on theta (non-adaptive forward process): the masks of :741-779 are replayed here
        on the CPU in fp32 only to know when the reference stops drawing; steps after that get zeros, which the
        kernels never use (every path is stopped)."""
        p, dt = self.problem, pt.tensor(self.delta_t_np)
        sq, B = pt.sqrt(dt), p.B.cpu()
        K = X.shape[0]
        X, stopped, xis = X.clone(), pt.zeros(K, dtype=pt.bool), []
        for n in range(self.N):
            xi = pt.randn(K, self.d)
            sel = ~stopped
            if int(sel.sum()) == 0:
                break
            xis.append(xi)
            X_prop = X + (pt.mm(B, xi.t()).t() * sq) * sel.float().unsqueeze(1)
            if p.boundary == 'sphere':
                new_sel = pt.sqrt(pt.sum(X ** 2, 1)) < p.boundary_distance
            elif p.boundary == 'two_spheres':
                r = pt.sqrt(pt.sum(X ** 2, 1))
                new_sel = (r > p.boundary_distance_1) & (r < p.boundary_distance_2)
            elif p.one_boundary:
                new_sel = pt.all(X_prop <= p.X_r, 1)
            else:
                new_sel = pt.all((X_prop >= p.X_l) & (X_prop <= p.X_r), 1)
            act = (new_sel & ~stopped).float().unsqueeze(1)
            X = X * (1 - act) + X_prop * act
            stopped = stopped | (~new_sel & ~stopped)
        xis += [pt.zeros(K, self.d)] * (self.N - len(xis))
        return pt.stack(xis)

    def initialize_training_data(self):
        """'inject': the reference's CPU draws in its order, pushed to the device.  'philox': start points from the
        device-side sampler / a per-iteration device generator, increments generated inside the kernels."""
        eng = self._get_engine()
        p, lo, hi = self.problem, self._k_lo, self._k_hi
        if self.noise == 'inject':
            Xb = self._sample_boundary_cpu()
            if p.boundary == 'sphere':
                X = pt.randn(self.K, self.d)
                X = p.boundary_distance * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * \
                    (pt.rand(self.K).unsqueeze(1) ** (1 / self.d))
            elif p.boundary == 'two_spheres':                  # solver.py:694-701: the batch shrinks to the annulus
                X = pt.randn(self.K_original, self.d)
                X = p.boundary_distance_2 * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * \
                    (pt.rand(self.K_original).unsqueeze(1) ** (1 / self.d))
                X = X[pt.sqrt(pt.sum(X ** 2, 1)) > p.boundary_distance_1, :]
                self.K = int(X.shape[0])
                lo, hi = 0, self.K
            else:
                X = (p.X_r - p.X_l) * pt.rand(self.K, self.d) + p.X_l
            xis = self._draw_increments_cpu(X)
            call = DiffusionCall(X[lo:hi].contiguous().to(self.device), None,
                                 xis[:, lo:hi].contiguous().to(self.device), self._iteration)
            call.Xb = Xb.to(self.device)
            call.gb = p.g(Xb).to(self.device)       # boundary data evaluated where the samples were drawn (the committor's
            return call                             # indicator |x| > a sits exactly on the inner sphere: rounding decides)
        gen = pt.Generator(device=self.device).manual_seed((self.seed * 1000003 + self._iteration) % (2 ** 63))
        if p.boundary == 'sphere':
            X0, _ = eng.sample(float(p.boundary_distance), self._iteration)
            Xb = pt.randn(self.K_boundary, self.d, device=self.device, generator=gen)
            Xb = p.boundary_distance * Xb / pt.sqrt(pt.sum(Xb ** 2, 1)).unsqueeze(1)
        elif p.boundary == 'two_spheres':
            X0, _ = eng.sample(float(p.boundary_distance_2), self._iteration)
            X0 = X0[pt.sqrt(pt.sum(X0 ** 2, 1)) > p.boundary_distance_1, :].contiguous()
            self.K = int(X0.shape[0])
            Xb = pt.randn(self.K_boundary, self.d, device=self.device, generator=gen)
            radii = pt.tensor([p.boundary_distance_1] * int(self.K_boundary / 2) +
                              [p.boundary_distance_2] * int(self.K_boundary / 2), device=self.device)
            Xb = radii.unsqueeze(1) * Xb / pt.sqrt(pt.sum(Xb ** 2, 1)).unsqueeze(1)
        else:
            Xb = (p.X_r - p.X_l) * pt.rand(self.K_boundary, self.d, device=self.device, generator=gen) + p.X_l
            face = pt.randint(0, self.d, (self.K_boundary,), device=self.device, generator=gen)
            side = pt.arange(self.K_boundary, device=self.device) >= self.K_boundary // 2
            Xb[pt.arange(self.K_boundary, device=self.device), face] = \
                pt.where(side | bool(p.one_boundary), pt.tensor(float(p.X_r), device=self.device),
                         pt.tensor(float(p.X_l), device=self.device))
            X0 = ((p.X_r - p.X_l) * pt.rand(self.K, self.d, device=self.device, generator=gen) + p.X_l)[lo:hi].contiguous()
        call = DiffusionCall(X0, None, None, self._iteration)
        call.Xb = Xb
        return call

    def gradient_descent(self, call):
        """fused rollout -> loss (solver.py:669-670, :790) -> fused backward -> Adam (:795-797)."""
        eng = self._get_engine()
        self._theta.grad.zero_()
        self._ensure_grad_views()
        V0, VE, Y = FusedDiffusion.apply(self._theta, eng, call, self.N)
        r = (VE - Y).double()
        ok = pt.isfinite(r)
        r = pt.where(ok, r, pt.zeros_like(r))
        sums = pt.stack([(r * r).sum().detach(), eng.stats_last[1], (~ok).sum().double(), eng.VL2_last.double().sum()])
        loss_local = self.alpha[0] * (r * r).sum() / self.K                                 # :790
        rank, _ = dist.world(self.process_group)
        lb = pt.zeros((), dtype=pt.float64, device=self.device)
        if self.boundary_loss and rank == 0:                                               # :669-670
            Xb = call.Xb.contiguous()
            Vb, _, _ = FusedDiffusion.apply(self._theta, eng, DiffusionCall(Xb, None, None, call.offset), 0)
            gb = call.gb if getattr(call, 'gb', None) is not None else self.problem.g(Xb)
            lb = self.alpha[1] * ((Vb.double() - gb.double()) ** 2).mean()
            loss_local = loss_local + lb
        loss_local.backward()
        self._ensure_grad_views()
        dist.all_reduce_sum_(self._theta.grad, self.process_group)
        sums = pt.cat([loss_local.detach().reshape(1), sums, lb.detach().reshape(1)])
        dist.all_reduce_sum_(sums, self.process_group)
        self.V.optim.step()
        return sums                       # [loss, sum r^2, K_count, #non-finite, sum V_L2, boundary loss]

    def train_step(self, l):
        t_0 = time.time()
        self._iteration = l
        call = self.initialize_training_data()
        loss, sr2, k_count, n_bad, vl2, lb = self.gradient_descent(call).tolist()   # the only host sync of the iteration
        self.loss_log.append(loss)
        self.loss_log_boundary.append(lb)
        self.loss_log_domain.append(loss - lb)
        self.K_log.append(int(k_count))                          # solver.py:775-776, :796
        self.V_L2_log.append(vl2 / self.K)                       # :733, :801
        if self.K_test_log is not None:                          # :802-806
            e = self.compute_test_error(self.K_test_log)
            self.V_test_L2.append(e[0])
            self.V_test_abs.append(e[1])
            self.V_test_rel_abs.append(e[2])
        t_1 = time.time()
        self.times.append(t_1 - t_0)
        self.path_steps_per_sec.append(self.K * self.N / max(t_1 - t_0, 1e-12))
        return loss

    def train(self):
        pt.manual_seed(self.seed)                                # solver.py:630-631
        np.random.seed(self.seed)
        for l in range(self.L):
            self.train_step(l)
            if self.verbose and l % self.print_every == 0:
                print('%d - loss = %.4e, v L2 error = %.4e, active: %d/%d, %.2f'
                      % (l, self.loss_log[-1], self.V_L2_log[-1], self.K_log[-1], self.K * self.N,
                         np.mean(self.times[-self.print_every:])))

    def compute_test_error(self, K_test, generator=None):
        """utilities.compute_test_error (:440-472), 'elliptic' modus, evaluated on the device: mean squared, mean
        absolute and mean relative error of V against problem.v_true on K_test fresh interior samples."""
        p = self.problem
        kw = dict(device=self.device, generator=generator)
        if p.boundary == 'sphere':
            X = pt.randn(K_test, self.d, **kw)
            X = p.boundary_distance * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * \
                (pt.rand(K_test, **kw).unsqueeze(1) ** (1 / self.d))
        else:
            X = (p.X_r - p.X_l) * pt.rand(K_test, self.d, **kw) + p.X_l
        with pt.no_grad():
            v_est = self.V(X).squeeze()
            v_true = p.v_true(X).squeeze()
            err = v_true - v_est
            return float((err ** 2).mean()), float(err.abs().mean()), float((err.abs() / v_true).mean())

    def V_L2_error(self, K_test=10000, seed=0):
        """rel. L2 error of V against problem.v_true on fresh interior samples (BASELINE metric)."""
        gen = pt.Generator(device=self.device).manual_seed(seed)
        p = self.problem
        if p.boundary == 'sphere':
            X = pt.randn(K_test, self.d, device=self.device, generator=gen)
            X = p.boundary_distance * X / pt.sqrt((X ** 2).sum(1, keepdim=True)) * \
                pt.rand(K_test, 1, device=self.device, generator=gen) ** (1 / self.d)
        else:
            X = (p.X_r - p.X_l) * pt.rand(K_test, self.d, device=self.device, generator=gen) + p.X_l
        with pt.no_grad():
            v = self.V(X).squeeze()
        ref = p.v_true(X)
        return float(pt.sqrt(((v - ref) ** 2).mean() / (ref ** 2).mean()))
