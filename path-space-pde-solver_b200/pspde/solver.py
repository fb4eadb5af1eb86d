"""Solver with the interface of the reference's ``Solver`` (solver.py:18-557), running the training hot path --
noise, N-step Euler-Maruyama rollout, per-step network evaluation, path-space loss and its gradient
(solver.py:433-499, :221) -- as fused sm_100a CUDA kernels (pspde.fused / libpspde.so).

Kept from the reference: the constructor signature (solver.py:20-25), the mutable attributes ``z_n``, ``y_0``,
``Phis`` with ``update_Phis()`` (:142-162), ``train()`` (:420), ``Z_n(X, t)`` (:360-362), the per-module Adam
(:198-200) and the result lists ``loss_log, u_L2_loss, Y_0_log, IS_rel_log, times, particles_close_to_target``
(:112-119).  Added keyword arguments:
    noise='philox' | 'inject'   in-kernel counter-based noise (default) or the reference's CPU draw
                                xi = randn(K, d, N+1) (:381) pushed to the device (parity mode: identical
                                torch RNG stream, hence identical trajectories to the reference)
    blowup_bound=1e6            trajectories with |Y_N - g(X_N)| >= blowup_bound are dropped from the batch like non-finite
                                ones (zero cotangent, counted in nonfinite_log); None / 0 switches it off.  At large K
                                the untrained control sends a few paths to |D| ~ 1e20: finite, but the loss becomes
                                1e34 and the gradient inf (the reference then trains on NaN)
    device                      CUDA device (default: current)
    process_group               torch.distributed group; K is the GLOBAL batch, sharded over its ranks
There is no CPU fallback; options of the reference that are off the fused path raise NotImplementedError.
"""
import json
import os
import time
from datetime import date

import numpy as np
import torch as pt

from . import _lib as L
from . import dist, losses
from .function_space import DenseNet, MySequential, SingleParam
from .fused import Call, FusedRollout, FusedRolloutAttached, FusedRolloutAttachedGeneral, RolloutEngine


class Solver:

    def __init__(self, name, problem, lr=0.001, L=10000, K=50, delta_t=0.05,
                 approx_method='control', loss_method='log-variance', time_approx='outer',
                 learn_Y_0=False, adaptive_forward_process=True, detach_forward=False,
                 early_stopping_time=10000, random_X_0=False, compute_gradient_variance=0,
                 IS_variance_K=0, IS_variance_iter=1, metastability_logs=None, print_every=100,
                 plot_trajectories=None, seed=42, save_results=False, u_l2_error_flag=True, log_gradient=False,
                 burgers_drift=False, verbose=True, noise='philox', device=None, process_group=None,
                 blowup_bound=1e6):
        self.problem, self.name = problem, name
        self.date = date.today().strftime('%Y-%m-%d')
        self.d, self.T, self.X_0 = problem.d, problem.T, problem.X_0
        self.Y_0 = pt.tensor([0.0])

        self.device = pt.device('cuda', pt.cuda.current_device()) if device is None else pt.device(device)
        self.seed = seed
        self.delta_t_np = delta_t
        self.delta_t = pt.tensor(self.delta_t_np).to(self.device)
        self.sq_delta_t = pt.sqrt(self.delta_t).to(self.device)
        self.N = int(np.floor(self.T / self.delta_t_np))         # host float64, like solver.py:41
        self.lr, self.L, self.K, self.random_X_0 = lr, L, K, random_X_0

        self.loss_method, self.approx_method, self.learn_Y_0 = loss_method, approx_method, learn_Y_0
        self.adaptive_forward_process, self.detach_forward = adaptive_forward_process, detach_forward
        self.early_stopping_time, self.burgers_drift = early_stopping_time, burgers_drift
        self.has_ref_solution = hasattr(problem, 'u_true')
        self.u_l2_error_flag = u_l2_error_flag and self.has_ref_solution
        if self.loss_method == 'relative_entropy':
            self.adaptive_forward_process = True                  # solver.py:61-62
        if self.loss_method == 'cross_entropy':
            self.learn_Y_0 = False                                # solver.py:63-64

        self.print_every, self.verbose, self.save_results = print_every, verbose, save_results
        self.compute_gradient_variance, self.IS_variance_K, self.IS_variance_iter = (
            compute_gradient_variance, IS_variance_K, IS_variance_iter)
        self.metastability_logs, self.plot_trajectories, self.log_gradient = (
            metastability_logs, plot_trajectories, log_gradient)
        self.noise, self.process_group = noise, process_group
        self.blowup_bound = blowup_bound
        if noise not in ('philox', 'inject'):
            raise ValueError("noise must be 'philox' or 'inject'")

        self.Phis, self.time_approx = [], time_approx
        pt.manual_seed(seed)                                      # solver.py:84
        if self.approx_method == 'control':
            self.y_0 = SingleParam(lr=self.lr).to(self.device)
            if self.time_approx == 'outer':
                self.z_n = [DenseNet(d_in=self.d, d_out=self.d, lr=self.lr, seed=seed) for _ in range(self.N)]
            elif self.time_approx == 'inner':
                self.z_n = MySequential(d_in=self.d + 1, d_out=self.d, lr=self.lr, seed=123)   # solver.py:91
            else:
                raise ValueError("time_approx must be 'outer' or 'inner'")
        else:
            raise NotImplementedError("approx_method=%r is off the fused hot path (only 'control')" % approx_method)
        self.update_Phis()
        for phi in self.Phis:
            phi.train()

        self.Y_0_log, self.loss_log, self.u_L2_loss, self.IS_rel_log = [], [], [], []
        self.times, self.grads_rel_error_log, self.particles_close_to_target = [], [], []
        self.path_steps_per_sec, self.nonfinite_log = [], []
        self._iteration = 0

    # ------------------------------------------------------------------ problem pass-throughs (solver.py:121-140)
    def b(self, x):
        return self.problem.b(x)

    def sigma(self, x):
        return self.problem.sigma(x)

    def h(self, t, x, y, z):
        return self.problem.h(t, x, y, z)

    def f(self, x, t):
        return self.problem.f(x, t)

    def g(self, x):
        return self.problem.g(x)

    def u_true(self, x, t):
        return self.problem.u_true(x, t)

    def v_true(self, x, t):
        return self.problem.v_true(x, t)

    # ------------------------------------------------------------------ parameters
    def _nets(self):
        return list(self.z_n) if self.time_approx == 'outer' else [self.z_n]

    def update_Phis(self):
        """Collect the trainable modules (solver.py:142-162) and re-home their parameters in ONE flat device buffer
        (theta layout of include/pspde.h: parameters() order, N stacked sets in 'outer' mode).  The nn.Parameters
        become views of that buffer and their .grad views of the flat gradient, so the kernels read theta and
        write dLoss/dtheta in place and every module's own Adam keeps working unchanged."""
        nets = self._nets()
        self.Phis = nets + [self.y_0] if self.learn_Y_0 else list(nets)
        for phi in self.Phis:
            phi.to(self.device)
        self.p = sum(int(np.prod(q.size())) for q in self.Phis[0].parameters() if q.requires_grad)
        spec = nets[0].net_spec() if hasattr(nets[0], 'net_spec') else None
        if spec is None:
            raise NotImplementedError("%s has no fused kernel (supported: DenseNet, MySequential)"
                                      % type(nets[0]).__name__)
        for m in nets[1:]:
            if m.net_spec() != spec:
                raise NotImplementedError("'outer' mode needs identically shaped networks")
        self._net_id, self._dims = spec
        params = [q for m in nets for q in m.parameters()]
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
        self._params = params
        self._engine = None
        self._flat_adam = None
        if self.log_gradient:
            self.gradient_log = pt.zeros(self.L, self.p)

    def zero_grad(self):
        self._theta.grad.zero_()
        if self.learn_Y_0 and self.y_0.Y_0.grad is not None:
            self.y_0.Y_0.grad.zero_()

    def optimization_step(self):
        """One Adam per module (solver.py:198-200).  In 'outer' mode that is N optimizers over N small networks; when
        they all are plain torch Adam with the same settings, ONE element-wise update of the flat buffer does the same
        arithmetic (the op sequence of torch.optim.Adam; bit-equal to its single-tensor form, within a few ulp of the update
        of its foreach CUDA kernels) in 7 kernels instead of 8 N.  The per-module optimizers keep
        owning the state -- their exp_avg / exp_avg_sq / step tensors are views of the flat state -- so reading them,
        stepping a module by hand or changing a learning rate (which drops back to the per-module loop) still works."""
        fa = self._flat_adam_state()
        rest = self.Phis
        if fa is not None:
            hyper = {self._adam_hyper(m.optim) for m in fa['nets']}
            steps = fa['step']
            if len(hyper) == 1 and None not in hyper and bool((steps == steps[0]).all()):
                lr, beta1, beta2, eps = next(iter(hyper))
                steps += 1
                t = float(steps[0])
                g, m1, m2 = self._theta.grad, fa['exp_avg'], fa['exp_avg_sq']
                if self._theta.is_cuda:                          # the same update in ONE launch (pspde_adam_flat)
                    import ctypes
                    lib = L.load()
                    vp = lambda x: ctypes.c_void_p(x.data_ptr())
                    with pt.cuda.device(self.device):
                        L.check(lib, lib.pspde_adam_flat(self._theta.numel(), vp(self._theta.data), vp(g), vp(m1), vp(m2), lr,
                                                         beta1, beta2, eps, int(t),
                                                         ctypes.c_void_p(pt.cuda.current_stream(self.device).cuda_stream)))
                else:
                    m1.lerp_(g, 1 - beta1)
                    m2.mul_(beta2).addcmul_(g, g, value=1 - beta2)
                    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
                    denom = m2.sqrt().div_(bc2 ** 0.5).add_(eps)
                    self._theta.data.addcdiv_(m1, denom, value=(lr / bc1) * -1)
                rest = [phi for phi in self.Phis if all(phi is not m for m in fa['nets'])]
        for phi in rest:
            phi.optim.step()

    @staticmethod
    def _adam_hyper(opt):
        """(lr, beta1, beta2, eps) of a plain Adam with one parameter group, else None."""
        if type(opt) is not pt.optim.Adam or len(opt.param_groups) != 1:
            return None
        g = opt.param_groups[0]
        if g.get('weight_decay', 0) != 0 or g.get('amsgrad') or g.get('maximize') or g.get('capturable') \
                or g.get('differentiable') or g.get('fused') or isinstance(g['lr'], pt.Tensor):
            return None
        return (float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']))

    def _flat_adam_state(self):
        """Flat Adam state shared with the per-module optimizers of the networks (built on first use after update_Phis);
        None when an optimizer is not a plain Adam over exactly its module's parameters (or, on the CPU, when there is only one
        network: torch's own step is then just as good)."""
        nets = self._nets()
        if self._flat_adam and self._flat_adam['opts'] != [id(m.optim) for m in nets]:
            self._flat_adam = None                                             # an optimizer was replaced: start over
        if self._flat_adam is None:
            self._flat_adam = False
            ok = (len(nets) > 1 or self._theta.is_cuda) and all(hasattr(m, 'optim') and self._adam_hyper(m.optim) is not None and
                                       [id(q) for q in m.optim.param_groups[0]['params']] == [id(q) for q in m.parameters()]
                                       for m in nets)
            if ok:
                m1, m2 = pt.zeros_like(self._theta.data), pt.zeros_like(self._theta.data)
                step = pt.zeros(len(self._params), dtype=pt.float32)          # torch keeps Adam's step counters on the host
                off, i = 0, 0
                for m in nets:
                    for q in m.parameters():
                        k, st = q.numel(), m.optim.state[q]
                        if len(st):                                            # already stepped by hand: adopt its state
                            m1[off:off + k].copy_(st['exp_avg'].reshape(-1))
                            m2[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
                            step[i] = float(st['step'])
                        st['step'], st['exp_avg'], st['exp_avg_sq'] = step[i], m1[off:off + k].view(q.shape), m2[off:off + k].view(q.shape)
                        off += k
                        i += 1
                self._flat_adam = dict(nets=nets, opts=[id(m.optim) for m in nets], exp_avg=m1, exp_avg_sq=m2, step=step)
        return self._flat_adam or None

    def _ensure_grad_views(self):
        # optimizers may have dropped .grad (zero_grad(set_to_none=True)); restore the views of the flat gradient
        off = 0
        g = self._theta.grad
        for q in self._params:
            k = q.numel()
            if q.grad is None or q.grad.data_ptr() != g.data_ptr() + 4 * off:
                q.grad = g[off:off + k].view(q.shape)
            off += k

    # ------------------------------------------------------------------ engine
    def _get_engine(self):
        if self._engine is None:
            unsupported = []
            if self.burgers_drift:
                unsupported.append('burgers_drift')
            if self.compute_gradient_variance:
                unsupported.append('compute_gradient_variance')
            if self.loss_method not in losses.SUPPORTED:
                unsupported.append('loss_method=%r' % self.loss_method)
            if unsupported:
                raise NotImplementedError("off the fused hot path: " + ", ".join(unsupported))
            rank, W = dist.world(self.process_group)
            lo, hi = dist.shard_range(self.K, rank, W)
            self._k_lo, self._k_hi = lo, hi
            tm = L.TIME_NONE if self.time_approx == 'outer' else L.TIME_FIRST
            self._engine = RolloutEngine(self.problem, self._net_id, self._dims, tm, hi - lo, self.N, self.delta_t_np,
                                         adaptive=self.adaptive_forward_process, k_offset=lo, K_global=self.K,
                                         seed=self.seed, device=self.device, blowup_bound=self.blowup_bound or 0.0)
            self._u_l2_on = False
            if self.u_l2_error_flag and hasattr(self.problem, 'u_true_table'):
                desc = self.problem.u_true_table(self.N, self.delta_t_np)
                if desc is not None:
                    self._engine.enable_u_l2(desc)
                    self._u_l2_on = True
            if self._engine.n_theta != self._theta.numel():
                raise RuntimeError("parameter count mismatch: modules %d vs kernel %d"
                                   % (self._theta.numel(), self._engine.n_theta))
        return self._engine

    def initialize_training_data(self):
        """Noise for one iteration.  'inject' reproduces solver.py:381 (CPU draw of the whole (K, d, N+1) tensor,
        then H2D); 'philox' draws nothing here -- the kernels generate the increments from (seed, iteration)."""
        eng = self._get_engine()
        if self.random_X_0:
            if self.noise == 'inject':                            # the reference's CPU draw, solver.py:366-367
                X0 = pt.randn(self.K, self.d)[self._k_lo:self._k_hi].to(self.device)
            else:
                # Philox mode: drawn on the device, one generator per (seed, iteration, shard start); K * d floats on the
                # host every iteration would be 400 MB at K = 2^20.  (Per-path starts therefore depend on the sharding,
                # unlike the in-kernel increments; the distribution does not.)
                gen = pt.Generator(device=self.device)
                gen.manual_seed((self.seed * 1000003 + self._iteration * 8191 + self._k_lo) % (2 ** 63))
                X0 = pt.randn(self._k_hi - self._k_lo, self.d, device=self.device, generator=gen)
            eng.set_x0(X0)
        xi = None
        if self.noise == 'inject':
            xi = pt.randn(self.K, self.d, self.N + 1)[self._k_lo:self._k_hi].to(self.device)
        return Call(offset=self._iteration, xi=xi)

    # ------------------------------------------------------------------ one training iteration
    def gradient_descent(self, call):
        """zero_grad -> fused rollout -> loss -> fused backward -> Adam (solver.py:202-223)."""
        eng = self._get_engine()
        self.zero_grad()
        self._ensure_grad_views()
        # X depends on theta only through an attached adaptive control (solver.py:451-469)
        attached = (not self.detach_forward) and self.adaptive_forward_process
        y0 = self.y_0.Y_0 if self.learn_Y_0 else None
        if (not attached) and self.loss_method in ('log-variance', 'moment'):
            # Detached forward, log-variance / moment loss (every log-variance notebook): forward launch -> statistics (one
            # all-reduce) -> ONE launch for the loss value and the per-path cotangents -> gradient launch(es), without going
            # through autograd and the dozen element-wise kernels of losses.value_and_cotangents (same arithmetic).
            import ctypes
            theta = self._theta.detach()
            y0c = None if y0 is None else y0.detach().contiguous()
            kept = eng.forward(theta, y0c, call, keep_rows=True)
            call.X_N, call.stats = eng.X_N, eng.stats
            stats = eng.stats
            if dist.world(self.process_group)[1] > 1:
                stats = dist.all_reduce_sum_(eng.stats.clone(), self.process_group)
            if getattr(self, '_wY', None) is None or self._wY.numel() != eng.K_local:
                self._wY = pt.empty(eng.K_local, dtype=pt.float32, device=self.device)
                self._lv_out = pt.empty(3, dtype=pt.float64, device=self.device)
            vp = lambda x: ctypes.c_void_p(x.data_ptr())
            with pt.cuda.device(self.device):
                L.check(eng.lib, eng.lib.pspde_lv_cotangents(eng.K_local, float(self.K), 1 if self.loss_method == 'moment' else 0,
                                                             vp(eng.Y_N), vp(eng.gX), vp(stats), vp(self._wY), vp(self._lv_out),
                                                             ctypes.c_void_p(pt.cuda.current_stream(self.device).cuda_stream)))
            if kept and eng.ckpt is not None:
                eng.grad_from_rows(theta, self._wY, call, self._theta.grad)
            else:
                eng.backward_detached(theta, self._wY, None, call, self._theta.grad)
            if self.learn_Y_0:
                self.y_0.Y_0.grad = self._wY.sum().reshape(1)
            loss, n_bad = self._lv_out[0].clone(), self._lv_out[1].clone()
        elif attached and self.loss_method == 'relative_entropy' and not self.learn_Y_0:
            loss_local = FusedRolloutAttached.apply(self._theta, eng, call)       # constant cotangents: one launch
            loss_local.backward()
            loss, n_bad = dist.all_reduce_sum_(pt.cat([loss_local.detach().double().reshape(1), call.stats[3:4]]),
                                               self.process_group)
        else:
            fn = FusedRolloutAttachedGeneral if attached else FusedRollout
            Y, gX, Zsum = fn.apply(self._theta, y0, eng, call)
            loss, wY, wZ, wG, n_bad = losses.value_and_cotangents(self.loss_method, Y.detach(), gX.detach(),
                                                                  Zsum.detach(), self.K, self.adaptive_forward_process,
                                                                  self.process_group, stats=call.stats)
            outs, cots = [], []
            for o, w in ((Y, wY), (Zsum, wZ)) + (((gX, wG),) if attached else ()):
                if w is not None:
                    outs.append(o); cots.append(w)
            pt.autograd.backward(outs, cots)
        self._ensure_grad_views()
        dist.all_reduce_sum_(self._theta.grad, self.process_group)
        if self.learn_Y_0 and self.y_0.Y_0.grad is not None:
            dist.all_reduce_sum_(self.y_0.Y_0.grad, self.process_group)
        self.optimization_step()
        u_l2 = pt.full((), float('nan'), dtype=pt.float64, device=self.device)
        if self._u_l2_on:
            u = eng.uL2.double()                               # filled by the forward launch of this iteration
            kept = pt.isfinite(u) & pt.isfinite(eng.Y_N)       # trajectories dropped from the batch carry Y_N = NaN
            u_l2 = dist.all_reduce_sum_(pt.where(kept, u, pt.zeros_like(u)).sum().reshape(1),
                                        self.process_group)[0] / self.K
        return pt.stack([loss.double(), n_bad.double(), u_l2])

    def train_step(self, l):
        """One full training iteration (noise -> rollout -> loss -> backward -> Adam -> logs), solver.py:431-531."""
        t_0 = time.time()
        self._iteration = l
        call = self.initialize_training_data()
        if self.learn_Y_0:
            self.Y_0_log.append(self.y_0.Y_0.item())
        res = self.gradient_descent(call)
        if self.log_gradient:
            self.gradient_log[l, :] = self._theta.grad[:self.p].detach().cpu()
        loss, n_bad, u_l2 = res.tolist()                      # the only host sync of the iteration (24 bytes D2H)
        self.loss_log.append(loss)
        self.nonfinite_log.append(int(n_bad))                 # trajectories dropped from the batch (blow-ups)
        self.u_L2_loss.append(u_l2)                           # mean_k sum_n |u - u*|^2 dt (solver.py:491-494, :515); NaN
                                                              # when the problem has no device table for u_true
        if self.metastability_logs is not None:
            target, epsilon = self.metastability_logs
            X = call.X_N
            close = (pt.sqrt(pt.sum((X - target) ** 2, 1)) < epsilon).float().sum().reshape(1).double()
            self.particles_close_to_target.append(
                (dist.all_reduce_sum_(close, self.process_group)[0] / self.K).item())
        if self.IS_variance_K > 0 and l % self.IS_variance_iter == 0:     # solver.py:521-528
            from .utilities import do_importance_sampling_me
            _, _, rel_IS = do_importance_sampling_me(self.problem, self, self.IS_variance_K)
            self.IS_rel_log.append(rel_IS)
        t_1 = time.time()
        self.times.append(t_1 - t_0)
        self.path_steps_per_sec.append(self.K * self.N / max(t_1 - t_0, 1e-12))
        return self.loss_log[-1]

    def train(self):
        pt.manual_seed(self.seed)                                 # solver.py:422
        if self.verbose:
            print('d = %d, L = %d, K = %d, delta_t = %.2e, lr = %.2e, %s, %s, %s, %s'
                  % (self.d, self.L, self.K, self.delta_t_np, self.lr, self.approx_method, self.time_approx,
                     self.loss_method, 'adaptive' if self.adaptive_forward_process else ''))
        for l in range(self.L):
            self.train_step(l)
            if self.verbose and l % self.print_every == 0:
                string = ('%d - loss: %.4e - u L2: %.4e - time/iter: %.2fs'
                          % (l, self.loss_log[-1], self.u_L2_loss[-1], np.mean(self.times[-self.print_every:])))
                if self.learn_Y_0:
                    string += ' - Y_0: %.4e' % self.Y_0_log[-1]
                if self.IS_variance_K > 0:
                    string += ' - rel IS: %.3e' % self.IS_rel_log[-1]
                print(string)
            if self.early_stopping_time is not None and l > self.early_stopping_time:
                if np.std(self.u_L2_loss[-self.early_stopping_time:]) / self.u_L2_loss[-1] < 0.02:
                    break
        if self.save_results is True:                             # solver.py:556-557
            self.save_logs()

    def save_logs(self, model_name='model'):
        """Run parameters, logs and the parameters of every trainable module as JSON under logs/ (solver.py:295-311: same
        keys and file naming; the modules stay on the device -- their flat-buffer views must not be re-homed)."""
        logs = dict(name=self.name, date=self.date, d=self.d, T=self.T, seed=self.seed, delta_t=self.delta_t_np, N=self.N,
                    lr=self.lr, K=self.K, loss_method=self.loss_method, learn_Y_0=self.learn_Y_0,
                    adaptive_forward_process=self.adaptive_forward_process, Y_0_log=self.Y_0_log, loss_log=self.loss_log,
                    u_L2_loss=self.u_L2_loss,
                    Phis_state_dict=[{k: v.detach().cpu().tolist() for k, v in z.state_dict().items()} for z in self.Phis])
        os.makedirs('logs', exist_ok=True)
        stem = 'logs/%s_%s_%s' % (model_name, self.name, self.date)
        path_name, i = stem + '.json', 1
        while os.path.isfile(path_name):
            i += 1
            path_name = '%s_%d.json' % (stem, i)
        with open(path_name, 'w') as f:
            json.dump(logs, f, indent=2)
        return path_name

    # ------------------------------------------------------------------ host-side evaluation of the control
    def Z_n_(self, X, n):
        if self.time_approx == 'outer':
            n = max(0, min(n, self.N - 1))
            return self.z_n[n](X)
        t_X = pt.cat([pt.ones([X.shape[0], 1], device=X.device) * n * self.delta_t, X], 1)
        return self.z_n(t_X)

    def Z_n(self, X, t):
        n = int(pt.ceil(pt.as_tensor(t, dtype=pt.float32, device=self.device) / self.delta_t))
        return self.Z_n_(X, n)

    def save_networks(self):
        path_name = 'output/%s_%s.pt' % (self.name, self.date)
        pt.save({'nn%d' % i: z.state_dict() for i, z in enumerate(self.Phis)}, path_name)
        print('\nnetworks data has been stored to file: %s' % path_name)

    def load_networks(self, cp_name):
        print('\nload network data from file: %s' % cp_name)
        checkpoint = pt.load(cp_name)
        for i, z in enumerate(self.Phis):
            z.load_state_dict(checkpoint['nn%d' % i])             # copies into the flat-buffer views in place
            z.eval()
