"""GeneralSolver with the interface of the reference's ``GeneralSolver`` (solver.py:934-1200) for the diffusion
loss on an unbounded domain (BASELINE config 4: HeatEquation, DenseNet value function): the sampling of the
initial points, the N-step rollout with the directional derivative of V, the loss and its gradient run as fused
sm_100a kernels (csrc/diffusion_kernels.cuh behind pspde_diffusion_* of include/pspde.h).

Kept from the reference: the constructor signature (:936-940), the attribute ``V`` (a DenseNet the caller may
replace before ``train()``), ``train()`` (:1001), the per-module Adam (``V.optim``, :1188) and the result lists
``loss_log, K_log, V_L2_log, times`` (:989-999).  Added keyword arguments as in ``pspde.Solver``:
    noise='philox' | 'inject'   in-kernel / device-side counter-based sampling (default) or the reference's CPU
                                draw order randn(K,d), rand(K), rand(K,1), N x randn(K,d) (:1045-1046, :1078, :1106)
    device, process_group       CUDA device; torch.distributed group (K is the GLOBAL batch, sharded over ranks)
Options off this path (other losses, bounded domains, adaptive / attached forward process, approx_method='Z')
raise NotImplementedError.  There is no CPU fallback.
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


class DiffusionCall:
    """Inputs of one iteration: device tensors X0 (K_local, d), t0 (K_local,), xis (N, K_local, d) or None."""

    def __init__(self, X0, t0, xis=None, offset=0):
        self.X0, self.t0, self.xis, self.offset = X0, t0, xis, int(offset)
        self.stats = None


class DiffusionEngine:
    """Device buffers + the library calls of the diffusion rollout for one (problem, network, K_local, N)."""

    def __init__(self, problem, dims, K_local, N, delta_t, k_offset=0, seed=42, device=None):
        self.lib = L.load()
        self.device = pt.device("cuda", pt.cuda.current_device()) if device is None else pt.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the fused rollout runs on CUDA devices only (got %s)" % self.device)
        if self.device.index is None:
            self.device = pt.device("cuda", pt.cuda.current_device())
        self._setup(problem, dims, K_local, N, delta_t, k_offset, seed)

    @on_own_device
    def _setup(self, problem, dims, K_local, N, delta_t, k_offset, seed):
        self.d, self.N, self.K_local, self.k_offset = int(problem.d), int(N), int(K_local), int(k_offset)
        self.dims, self.seed, self.T = list(dims), int(seed), float(getattr(problem, 'T', 1.0))
        self.dt = float(pt.tensor(delta_t, dtype=pt.float32))
        pid, flags, pack = problem.functor_pack()
        self.problem_id, self.flags = pid, flags
        self.pack = pack.to(self.device)
        f32 = dict(dtype=pt.float32, device=self.device)
        self.V0, self.VE, self.Y = (pt.empty(self.K_local, **f32) for _ in range(3))
        self.X_end = pt.empty(self.K_local, self.d, **f32)
        self.t_end = pt.empty(self.K_local, **f32)
        self.stats = pt.zeros(4, dtype=pt.float64, device=self.device)
        cfg = self.cfg(self.K_local, self.N, None, 0)
        self.n_theta = int(self.lib.pspde_theta_size(ctypes.byref(cfg)))
        nbytes = self.workspace_bytes(cfg)
        if self.n_theta < 0 or nbytes == 0:
            raise RuntimeError("libpspde: %s" % self.lib.pspde_last_error().decode())
        self.workspace = pt.empty(nbytes, dtype=pt.uint8, device=self.device)

    def workspace_bytes(self, cfg):
        return int(self.lib.pspde_diffusion_workspace_bytes(ctypes.byref(cfg), ctypes.c_float(self.T)))

    def cfg(self, K, N, xis, offset):
        noise, strides = L.NOISE_PHILOX, (0, 0, 0)
        if xis is not None:
            if xis.dtype != pt.float32 or xis.device != self.device or tuple(xis.shape) != (N, K, self.d):
                raise ValueError("xis must be a float32 (N, K_local, d) tensor on %s" % self.device)
            noise, strides = L.NOISE_INJECT, (xis.stride(1), xis.stride(2), xis.stride(0))
        return L.make_cfg(K, self.d, N, self.dt, self.problem_id, L.NET_DENSENET, self.dims, L.TIME_LAST,
                          adaptive=False, k_offset=self.k_offset, problem_flags=self.flags, noise_mode=noise,
                          seed=self.seed, offset=offset, xi_strides=strides)

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _stream(self):
        return ctypes.c_void_p(pt.cuda.current_stream(self.device).cuda_stream)

    @on_own_device
    def sample(self, radius, offset):
        """X_0 uniform in the ball, t_0 uniform in [0, T) from Philox (solver.py:1045-1046, :1078)."""
        X0 = pt.empty(self.K_local, self.d, dtype=pt.float32, device=self.device)
        t0 = pt.empty(self.K_local, dtype=pt.float32, device=self.device)
        cfg = self.cfg(self.K_local, self.N, None, offset)
        L.check(self.lib, self.lib.pspde_diffusion_sample(ctypes.byref(cfg), ctypes.c_float(radius),
                                                          ctypes.c_float(self.T), self._p(X0), self._p(t0),
                                                          self._stream()))
        return X0, t0

    @on_own_device
    def forward(self, theta, X0, t0, xis, offset, N=None, outs=None):
        K = X0.shape[0]
        N = self.N if N is None else N
        cfg = self.cfg(K, N, xis if N > 0 else None, offset)
        V0, VE, Y, X_end, t_end, stats = outs if outs is not None else (self.V0, self.VE, self.Y, self.X_end,
                                                                        self.t_end, self.stats)
        rc = self.lib.pspde_diffusion_fwd(ctypes.byref(cfg), ctypes.c_float(self.T), self._p(theta), self._p(self.pack),
                                          self._p(X0), self._p(t0), self._p(xis if N > 0 else None), self._p(V0),
                                          self._p(VE), self._p(Y), self._p(X_end), self._p(t_end), self._p(stats),
                                          self._p(self.workspace), self.workspace.numel(), self._stream())
        L.check(self.lib, rc)

    @on_own_device
    def backward(self, theta, X0, t0, xis, offset, c0, cE, cD, grad_out, N=None):
        K = X0.shape[0]
        N = self.N if N is None else N
        cfg = self.cfg(K, N, xis if N > 0 else None, offset)
        rc = self.lib.pspde_diffusion_bwd(ctypes.byref(cfg), ctypes.c_float(self.T), self._p(theta), self._p(self.pack),
                                          self._p(X0), self._p(t0), self._p(xis if N > 0 else None), self._p(c0),
                                          self._p(cE), self._p(cD), self._p(grad_out), self._p(self.workspace),
                                          self.workspace.numel(), self._stream())
        L.check(self.lib, rc)


class FusedDiffusion(pt.autograd.Function):
    """theta -> per-path (V(X_0,t_0), V(X_end,t_end), Y_end) of the diffusion rollout (solver.py:1076-1163).
    With N = 0 it is a plain batched evaluation of V (used for the terminal-condition term, :1063-1064)."""

    @staticmethod
    def forward(ctx, theta, engine, call, N):
        theta_c = theta.detach().contiguous()
        K = call.X0.shape[0]
        if N == engine.N and K == engine.K_local:
            outs = None
            engine.forward(theta_c, call.X0, call.t0, call.xis, call.offset)
            V0, VE, Y = engine.V0.clone(), engine.VE.clone(), engine.Y.clone()
            call.stats = engine.stats
        else:
            f32 = dict(dtype=pt.float32, device=engine.device)
            V0, VE, Y = (pt.empty(K, **f32) for _ in range(3))
            outs = (V0, VE, Y, None, None, None)
            engine.forward(theta_c, call.X0, call.t0, call.xis, call.offset, N=N, outs=outs)
        ctx.engine, ctx.call, ctx.N = engine, call, N
        ctx.save_for_backward(theta_c)
        return V0, VE, Y

    @staticmethod
    def backward(ctx, gV0, gVE, gY):
        (theta_c,) = ctx.saved_tensors
        engine, call = ctx.engine, ctx.call
        z = lambda t: None if t is None else t.contiguous().float()
        gV0, gVE, gY = z(gV0), z(gVE), z(gY)
        # Y_end = V0 + sum_n (directional derivative)_n: its cotangent reaches V0 and every active step
        c0 = gY if gV0 is None else (gV0 if gY is None else gV0 + gY)
        grad = pt.empty(engine.n_theta, dtype=pt.float32, device=engine.device)
        engine.backward(theta_c, call.X0, call.t0, call.xis, call.offset, c0, gVE, gY, grad, N=ctx.N)
        return grad, None, None, None


class GeneralSolver:

    def __init__(self, problem, name, seed=42, delta_t=0.01, N=50, lr=0.001, L=100000, K=200, K_boundary=50,
                 alpha=[1.0, 1.0, 1.0], adaptive_forward_process=False, detach_forward=True, print_every=100,
                 verbose=True, approx_method='Y', sample_center=False, loss_method='diffusion',
                 loss_with_stopped=False, K_test_log=None, PINN_log_variance=False, log_loss_parts=False,
                 boundary_loss=True, full_hessian=False, uniform_square=False, solve_linear_L2_projection=False,
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
        self.adaptive_forward_process, self.detach_forward = adaptive_forward_process, detach_forward
        self.approx_method, self.sample_center, self.loss_method = approx_method, sample_center, loss_method
        self.loss_with_stopped, self.boundary_loss = loss_with_stopped, boundary_loss
        self.print_every, self.verbose = print_every, verbose
        self.noise, self.process_group = noise, process_group
        self.uniform_square = uniform_square
        off = []
        if loss_method != 'diffusion':
            off.append('loss_method=%r' % loss_method)
        if approx_method != 'Y':
            off.append('approx_method=%r' % approx_method)
        if adaptive_forward_process or not detach_forward:
            off.append('adaptive / attached forward process')
        if getattr(problem, 'boundary', 'unbounded') != 'unbounded':
            off.append('boundary=%r' % problem.boundary)
        if sample_center or loss_with_stopped or solve_linear_L2_projection or full_hessian \
                or PINN_log_variance or K_test_log is not None:
            off.append('sample_center / loss_with_stopped / L2 projection / K_test_log')
        if off:
            raise NotImplementedError("off the fused hot path: " + ", ".join(off))
        if noise not in ('philox', 'inject'):
            raise ValueError("noise must be 'philox' or 'inject'")
        pt.manual_seed(seed)                                      # solver.py:975
        self.V = DenseNet(d_in=self.d + 1, d_out=1, lr=self.lr, seed=seed).to(self.device)   # :977
        self.K_test_log = K_test_log
        self.Y_0_log, self.loss_log, self.loss_log_domain, self.loss_log_boundary = [], [], [], []
        self.u_L2_log, self.V_L2_log, self.V_test_L2, self.V_test_abs, self.V_test_rel_abs = [], [], [], [], []
        self.times, self.lambda_log, self.K_log, self.path_steps_per_sec = [], [], [], []
        self.log_loss_parts = log_loss_parts
        self._engine, self._V_homed, self._iteration = None, None, 0

    # ------------------------------------------------------------------ parameters: one flat device buffer
    def _home_parameters(self):
        """Re-home the parameters of self.V in one flat buffer (theta layout of include/pspde.h); .grad become views
        of the flat gradient so that V.optim (Adam, solver.py:1188) works unchanged."""
        if self._V_homed is self.V:
            return
        self.V.to(self.device)
        params = list(self.V.parameters())
        n = sum(q.numel() for q in params)
        flat = pt.empty(n, dtype=pt.float32, device=self.device)
        gflat = pt.zeros(n, dtype=pt.float32, device=self.device)
        off = 0
        for q in params:
            k = q.numel()
            flat[off:off + k].copy_(q.data.reshape(-1))
            q.data = flat[off:off + k].view(q.shape)
            q.grad = gflat[off:off + k].view(q.shape)
            off += k
        self._theta = flat.requires_grad_(True)
        self._theta.grad = gflat
        self._params, self._V_homed, self._engine = params, self.V, None

    def _ensure_grad_views(self):
        off, g = 0, self._theta.grad
        for q in self._params:
            k = q.numel()
            if q.grad is None or q.grad.data_ptr() != g.data_ptr() + 4 * off:
                q.grad = g[off:off + k].view(q.shape)
            off += k

    def _get_engine(self):
        self._home_parameters()
        if self._engine is None:
            net_id, dims = self.V.net_spec()
            if net_id != L.NET_DENSENET or dims[0] != self.d + 1 or dims[-1] != 1:
                raise NotImplementedError("the diffusion kernels need V = DenseNet(d_in=d+1, d_out=1)")
            rank, W = dist.world(self.process_group)
            self._k_lo, self._k_hi = dist.shard_range(self.K, rank, W)
            self._engine = DiffusionEngine(self.problem, dims, self._k_hi - self._k_lo, self.N, self.delta_t_np,
                                           k_offset=self._k_lo, seed=self.seed, device=self.device)
            if self._engine.n_theta != self._theta.numel():
                raise RuntimeError("parameter count mismatch: module %d vs kernel %d"
                                   % (self._theta.numel(), self._engine.n_theta))
        return self._engine

    # ------------------------------------------------------------------ one iteration
    def initialize_training_data(self):
        """'inject': the reference's CPU draws in its order, pushed to the device.  'philox': X_0, t_0 from the
        device-side sampler, increments generated inside the kernels."""
        eng = self._get_engine()
        lo, hi = self._k_lo, self._k_hi
        if self.noise == 'inject':
            R = self.problem.boundary_distance
            if self.uniform_square:                                                         # solver.py:1041-1043
                X = pt.rand(self.K, self.d) * 2 - 1
                X = R * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * (pt.rand(self.K).unsqueeze(1))
            else:
                X = pt.randn(self.K, self.d)
                X = R * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * (pt.rand(self.K).unsqueeze(1) ** (1 / self.d))
            t0 = pt.rand(self.K, 1) * self.problem.T
            xis = self._draw_increments_cpu(t0)
            return DiffusionCall(X[lo:hi].contiguous().to(self.device), t0[lo:hi, 0].contiguous().to(self.device),
                                 xis[:, lo:hi].contiguous().to(self.device), self._iteration)
        X0, t0 = eng.sample(float(self.problem.boundary_distance), self._iteration)
        if self.uniform_square:      # direction from the cube, radius U (not U^(1/d)): a per-iteration device generator
            gen = pt.Generator(device=self.device).manual_seed((self.seed * 1000003 + self._iteration) % (2 ** 63))
            X = pt.rand(self.K, self.d, device=self.device, generator=gen) * 2 - 1
            X = float(self.problem.boundary_distance) * X / pt.sqrt(pt.sum(X ** 2, 1)).unsqueeze(1) * \
                pt.rand(self.K, 1, device=self.device, generator=gen)
            X0 = X[lo:hi].contiguous()
        return DiffusionCall(X0, t0, None, self._iteration)

    def _draw_increments_cpu(self, t0):
        """'inject' only: one randn(K, d) per step AFTER the all-stopped check (solver.py:1093-1094 break before :1106), so
        that a whole training loop consumes the CPU RNG stream exactly like the reference also when N * delta_t > T (every
        path stops early).  On the unbounded domain the stop mask is the time test of :1131 alone; it is replayed here in
        fp32 like the reference's (K, 1) tensor t_n.  Steps after the break get zeros, which the kernels never use."""
        dt, T = pt.tensor(self.delta_t_np), self.problem.T
        t = t0.squeeze(1).clone()
        stopped, xis = pt.zeros(self.K, dtype=pt.bool), []
        for n in range(self.N):
            if int((~stopped).sum()) == 0:
                break
            xis.append(pt.randn(self.K, self.d))
            new_sel = (t + dt) <= T
            t = t + dt * (new_sel & ~stopped).float()
            stopped = stopped | (~new_sel & ~stopped)
        xis += [pt.zeros(self.K, self.d)] * (self.N - len(xis))
        return pt.stack(xis)

    def gradient_descent(self, call):
        """fused rollout -> loss (solver.py:1063-1064, :1163) -> fused backward -> Adam (:1187-1188)."""
        eng = self._get_engine()
        self._theta.grad.zero_()
        self._ensure_grad_views()
        V0, VE, Y = FusedDiffusion.apply(self._theta, eng, call, self.N)
        r = (VE - Y).double()
        ok = pt.isfinite(r)
        r = pt.where(ok, r, pt.zeros_like(r))
        sums = pt.stack([(r * r).sum().detach(), call.stats[1], (~ok).sum().double()])
        loss_local = self.alpha[0] * (r * r).sum() / self.K                                 # :1163
        # terminal-condition term on the first K_boundary points of the GLOBAL batch (:1063-1064): every rank takes its
        # slice of that prefix and contributes sum / K_boundary, so loss and gradient do not depend on the number of ranks
        Kb = min(self.K_boundary, self.K)
        nb = max(0, min(Kb, self._k_hi) - self._k_lo)
        if self.boundary_loss and nb > 0:
            Xb = call.X0[:nb].contiguous()
            tb = pt.full((nb,), float(self.problem.T), dtype=pt.float32, device=self.device)
            Vb, _, _ = FusedDiffusion.apply(self._theta, eng, DiffusionCall(Xb, tb, None, call.offset), 0)
            lb = self.alpha[1] * ((Vb.double() - self.problem.f(Xb).double()) ** 2).sum() / Kb
            loss_local = loss_local + lb
        loss_local.backward()
        self._ensure_grad_views()
        dist.all_reduce_sum_(self._theta.grad, self.process_group)
        sums = pt.cat([loss_local.detach().reshape(1), sums])
        dist.all_reduce_sum_(sums, self.process_group)
        self.V.optim.step()
        return sums                                              # [loss, sum r^2, K_count, #non-finite]

    def train_step(self, l):
        t_0 = time.time()
        self._iteration = l
        call = self.initialize_training_data()
        loss, _, k_count, n_bad = self.gradient_descent(call).tolist()     # the only host sync of the iteration
        self.loss_log.append(loss)
        self.K_log.append(int(k_count))                          # solver.py:1151-1152, :1168
        self.V_L2_log.append(0.0)                                # the reference logs mean(V_L2) = 0 (:1110 is commented out)
        t_1 = time.time()
        self.times.append(t_1 - t_0)
        self.path_steps_per_sec.append(self.K * self.N / max(t_1 - t_0, 1e-12))
        return loss

    def train(self):
        pt.manual_seed(self.seed)                                # solver.py:1003
        for l in range(self.L):
            self.train_step(l)
            if self.verbose and l % self.print_every == 0:
                print('%d - loss = %.4e, v L2 error = %.4e, active: %d/%d, %.2f'
                      % (l, self.loss_log[-1], self.V_L2_log[-1], self.K_log[-1], self.K * self.N,
                         np.mean(self.times[-self.print_every:])))

    def V_L2_error(self, K_test=10000, seed=0):
        """rel. L2 error of V(., 0) against problem.v_true on x ~ Unif(ball) (BASELINE metric)."""
        gen = pt.Generator().manual_seed(seed)
        X = pt.randn(K_test, self.d, generator=gen)
        X = self.problem.boundary_distance * X / pt.sqrt((X ** 2).sum(1, keepdim=True)) * \
            pt.rand(K_test, 1, generator=gen) ** (1 / self.d)
        X = X.to(self.device)
        with pt.no_grad():
            v = self.V(pt.cat([X, pt.zeros(K_test, 1, device=self.device)], 1)).squeeze()
        ref = self.problem.v_true(X, 0.0)
        return float(pt.sqrt(((v - ref) ** 2).mean() / (ref ** 2).mean()))
