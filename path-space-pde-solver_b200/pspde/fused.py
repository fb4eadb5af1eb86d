"""torch.autograd bridge to the fused CUDA rollout (C ABI: include/pspde.h).

``RolloutEngine`` owns what one Solver needs on the device (problem functor pack, x0, workspace, per-path output
buffers) and exposes the library calls with torch tensors.  ``FusedRollout`` / ``FusedRolloutAttached`` are the
autograd Functions that replace the body of the reference's training iteration (solver.py:433-499 forward,
:221 backward):

  detached forward (detach_forward=True)      Y_N, gX, Zsum = FusedRollout.apply(theta, y0, engine, call)
      backward: per-path cotangents (dL/dY_N, dL/dZsum) -> dL/dtheta, dL/dy0.  When the forward pass could keep its
      operand rows (tensor-core shape class, adaptive process, buffer affordable) and dL/dZsum is absent, one gradient
      launch over those rows; else the checkpointed / recompute backward rollout.
  attached forward (relative entropy loss)     loss = FusedRolloutAttached.apply(theta, engine, call)
      forward and adjoint run in one kernel; backward hands the stored gradient back.

There is no CPU fallback: constructing an engine without the CUDA library raises.
"""
import ctypes
import functools
import os

import torch as pt

from . import _lib as L


class Call:
    """What varies from one training iteration to the next."""

    def __init__(self, offset=0, xi=None):
        self.offset = int(offset)    # Philox stream id (iteration counter)
        self.xi = xi                 # injected increments, reference layout (K_local, d, N+1) on the device, or None
        self.X_N = None              # filled by the forward pass
        self.stats = None
        self.grad_enabled = True     # set by FusedRollout.apply: was autograd recording when the rollout was called


ROWS_BUFFER_DEFAULT_GB = 8.0        # PSPDE_FWD_CKPT_MAX_GB: opt in to more (the buffer scales with K * N)
ROWS_BUFFER_FREE_FRACTION = 0.25    # never more than a quarter of what is free on the device right now


def rows_buffer_bytes(need, n_tiles, cap, free):
    """Bytes of the single-rollout step's row buffer: whole 128-path tiles, at most `need` (all tiles), `cap`
    (PSPDE_FWD_CKPT_MAX_GB, default 8 GB -- C2's 7.1 GB fit, C3 / C5 take the K-independent wave-checkpointed backward
    unless the user opts in) and a quarter of the free device memory (another solver on the same GPU keeps its room);
    0 when that is less than a tenth of the batch (not worth keeping a second code path busy)."""
    tile = need // n_tiles
    take = int(min(need, cap, ROWS_BUFFER_FREE_FRACTION * free)) // tile * tile
    return take if take >= max(tile, need // 10) else 0


def on_own_device(method):
    """libpspde launches on the CURRENT CUDA device and sizes its grids from it (api_common.h): make the engine's device
    current for the duration of the call, so that device='cuda:1' works while another device is current."""
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        with pt.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapped


class RolloutEngine:
    def __init__(self, problem, net_id, dims, time_mode, K_local, N, delta_t, adaptive=True, k_offset=0,
                 K_global=None, seed=42, device=None, want_X_N=True, blowup_bound=0.0):
        self.lib = L.load()
        self.blowup_bound = float(blowup_bound)     # pspde_cfg::d_abs_max of the training rollouts (0: off)
        self.device = pt.device("cuda", pt.cuda.current_device()) if device is None else pt.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the fused rollout runs on CUDA devices only (got %s)" % self.device)
        if self.device.index is None:
            self.device = pt.device("cuda", pt.cuda.current_device())
        self._setup(problem, net_id, dims, time_mode, K_local, N, delta_t, adaptive, k_offset, K_global, seed, want_X_N)

    @on_own_device
    def _setup(self, problem, net_id, dims, time_mode, K_local, N, delta_t, adaptive, k_offset, K_global, seed, want_X_N):
        self.d, self.N, self.K_local, self.k_offset = int(problem.d), int(N), int(K_local), int(k_offset)
        self.K_global = int(K_global if K_global is not None else K_local)
        self.net_id, self.dims, self.time_mode = net_id, list(dims), time_mode
        self.adaptive, self.seed = bool(adaptive), int(seed)
        # delta_t enters the arithmetic as an fp32 scalar, like the reference's 0-dim tensor (solver.py:39)
        self.dt = float(pt.tensor(delta_t, dtype=pt.float32))
        pid, flags, pack = problem.functor_pack()
        self.problem_id, self.flags = pid, flags
        self.pack = pack.to(self.device)
        self.x0 = problem.X_0.detach().to(self.device, pt.float32).contiguous()
        self.x0_per_path = False
        self.want_X_N = want_X_N
        f32 = dict(dtype=pt.float32, device=self.device)
        self.Y_N, self.gX, self.Zsum = (pt.empty(self.K_local, **f32) for _ in range(3))
        self.X_N = pt.empty(self.K_local, self.d, **f32) if want_X_N else None
        self.stats = pt.zeros(4, dtype=pt.float64, device=self.device)
        cfg = self.cfg(Call())
        self.n_theta = int(self.lib.pspde_theta_size(ctypes.byref(cfg)))
        if self.n_theta < 0:
            raise RuntimeError("libpspde: %s" % self.lib.pspde_last_error().decode())
        nbytes = int(self.lib.pspde_workspace_bytes(ctypes.byref(cfg)))
        if nbytes == 0:
            raise RuntimeError("libpspde: %s" % self.lib.pspde_last_error().decode())
        self.workspace = pt.empty(nbytes, dtype=pt.uint8, device=self.device)
        self.udiag, self.uL2 = None, None
        # Single-rollout step (pspde_rollout_fwd_ckpt): the training forward keeps the operand rows of its tiles for the
        # gradient kernel.  All tiles need K_local * N * ~1 KB (C2: 7.1 GB, C5: 228 GB), i.e. a K x N x d tape, so this is
        # bounded: the buffer takes what fits below PSPDE_FWD_CKPT_MAX_GB (default 8) and a quarter of the free device
        # memory, and nothing when that is less than a tenth of the batch -- the tiles it does not hold go through the
        # K-independent wave-checkpointed backward -- and it is dropped for good by the first backward that carries a
        # cotangent on Z_sum (those losses need the rollout with the cotangents in hand).
        self.ckpt, self.ckpt_ok, self._ckpt_need = None, True, 0
        self.rows_serial = 0         # row-keeping forwards so far: a backward may use the rows only if they are its own forward's

    def _fwd_ckpt_buffer(self, cfg):
        """The forward checkpoint buffer (whole tiles), or None if this configuration / device cannot take it."""
        if not self.ckpt_ok:
            return None
        need = int(self.lib.pspde_fwd_ckpt_bytes(ctypes.byref(cfg)))
        if need == 0:
            return None
        if self.ckpt is None or self._ckpt_need != need:
            self.ckpt = None
            cap = float(os.environ.get("PSPDE_FWD_CKPT_MAX_GB", str(ROWS_BUFFER_DEFAULT_GB))) * 2 ** 30
            free, _ = pt.cuda.mem_get_info(self.device)
            take = rows_buffer_bytes(need, (self.K_local + 127) // 128, cap, free)
            if take == 0:
                self.ckpt_ok = False
                return None
            try:
                self.ckpt, self._ckpt_need = pt.empty(take, dtype=pt.uint8, device=self.device), need
            except pt.cuda.OutOfMemoryError:                  # fragmented pool: stay on the two-rollout step
                self.ckpt_ok = False
                return None
        return self.ckpt

    def set_x0(self, x0):
        """(d,) broadcast start or (K_local, d) per-path starts (random_X_0, solver.py:366-367)."""
        x0 = x0.detach().to(self.device, pt.float32).contiguous()
        self.x0, self.x0_per_path = x0, x0.dim() == 2

    def cfg(self, call):
        noise = L.NOISE_PHILOX if call.xi is None else L.NOISE_INJECT
        strides = (0, 0, 0)
        if call.xi is not None:
            xi = call.xi
            if xi.dtype != pt.float32 or xi.device != self.device or tuple(xi.shape) != (self.K_local, self.d, self.N + 1):
                raise ValueError("xi must be a float32 (K_local, d, N+1) tensor on %s" % self.device)
            strides = (xi.stride(0), xi.stride(1), xi.stride(2))
        return L.make_cfg(self.K_local, self.d, self.N, self.dt, self.problem_id, self.net_id, self.dims,
                          self.time_mode, adaptive=self.adaptive, k_offset=self.k_offset, problem_flags=self.flags,
                          noise_mode=noise, seed=self.seed, offset=call.offset, x0_per_path=self.x0_per_path,
                          xi_strides=strides, d_abs_max=self.blowup_bound)

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _xi_ptr(self, call):
        # slice n+1 of the reference layout drives step n (solver.py:472)
        return None if call.xi is None else ctypes.c_void_p(call.xi.data_ptr() + 4 * call.xi.stride(2))

    def _stream(self):
        return ctypes.c_void_p(pt.cuda.current_stream(self.device).cuda_stream)

    def enable_u_l2(self, desc):
        """desc = problem.u_true_table(N, delta_t): switches on the u_L2 diagnostic of solver.py:491-494."""
        self._utab = desc["table"].to(self.device, pt.float32).contiguous()
        self.uL2 = pt.zeros(self.K_local, dtype=pt.float32, device=self.device)
        u = L.pspde_udiag()
        u.mode, u.nx1, u.d1 = int(desc["mode"]), int(desc.get("nx1", 0)), int(desc.get("d1", 0))
        u.xb, u.dx = float(desc.get("xb", 0.0)), float(desc.get("dx", 0.0))
        u.table, u.uL2 = self._utab.data_ptr(), self.uL2.data_ptr()
        u.quirk_path = self.K_global - 1 if desc.get("quirk_last") else -1
        self.udiag = u

    @on_own_device
    def forward(self, theta, y0, call, keep_rows=False):
        """keep_rows: training forward -- returns True if the operand rows were kept for `grad_from_rows`."""
        cfg = self.cfg(call)
        diag = None if self.udiag is None else ctypes.byref(self.udiag)
        ck = self._fwd_ckpt_buffer(cfg) if keep_rows else None
        rc = self.lib.pspde_rollout_fwd_ckpt(ctypes.byref(cfg), self._p(theta), self._p(self.pack), self._p(self.x0),
                                             self._p(y0), self._xi_ptr(call), self._p(self.X_N), self._p(self.Y_N),
                                             self._p(self.gX), self._p(self.Zsum), self._p(self.stats), diag,
                                             self._p(ck), 0 if ck is None else ck.numel(),
                                             self._p(self.workspace), self.workspace.numel(), self._stream())
        L.check(self.lib, rc)
        if ck is not None:
            self.rows_serial += 1       # the buffer now holds THIS forward's rows
        return ck is not None

    @on_own_device
    def grad_from_rows(self, theta, wY, call, grad_out):
        """dL/dtheta from the rows the last training forward kept (pspde_grad_from_fwd_ckpt)."""
        cfg = self.cfg(call)
        rc = self.lib.pspde_grad_from_fwd_ckpt(ctypes.byref(cfg), self._p(theta), self._p(self.pack), self._p(self.x0),
                                               self._xi_ptr(call), self._p(self.ckpt), self.ckpt.numel(),
                                               self._p(wY), self._p(grad_out), self._p(self.workspace),
                                               self.workspace.numel(), self._stream())
        L.check(self.lib, rc)

    @on_own_device
    def backward_detached(self, theta, wY, wZ, call, grad_out):
        cfg = self.cfg(call)
        rc = self.lib.pspde_rollout_bwd_detached(ctypes.byref(cfg), self._p(theta), self._p(self.pack),
                                                 self._p(self.x0), self._xi_ptr(call), self._p(wY), self._p(wZ),
                                                 self._p(grad_out), self._p(self.workspace), self.workspace.numel(),
                                                 self._stream())
        L.check(self.lib, rc)

    @on_own_device
    def attached(self, theta, call, grad_out, y0=None, wY=None, wZ=None, wG=None):
        """wY = wZ = wG = None: relative entropy in one launch (constant cotangents 1 / K_global)."""
        cfg = self.cfg(call)
        # the u_L2 diagnostic belongs to the forward sweep; the two-phase form (per-path cotangents) already got it from
        # its forward launch
        diag = None if (self.udiag is None or wY is not None or wZ is not None or wG is not None) else ctypes.byref(self.udiag)
        rc = self.lib.pspde_rollout_attached_diag(ctypes.byref(cfg), self._p(theta), self._p(self.pack), self._p(self.x0),
                                                  self._p(y0), self._xi_ptr(call), ctypes.c_float(1.0 / self.K_global),
                                                  self._p(wY), self._p(wZ), self._p(wG), self._p(self.X_N),
                                                  self._p(self.Y_N), self._p(self.gX), self._p(self.Zsum),
                                                  self._p(self.stats), diag, self._p(grad_out), self._p(self.workspace),
                                                  self.workspace.numel(), self._stream())
        L.check(self.lib, rc)

    @on_own_device
    def philox_dump(self, offset=0):
        """Increments the kernels generate for iteration `offset`, in the reference layout (K_local, d, N+1)."""
        out = pt.empty(self.N, self.K_local, self.d, dtype=pt.float32, device=self.device)
        cfg = self.cfg(Call(offset))
        L.check(self.lib, self.lib.pspde_philox_dump(ctypes.byref(cfg), self._p(out), self._stream()))
        xi = pt.zeros(self.K_local, self.d, self.N + 1, dtype=pt.float32, device=self.device)
        xi[:, :, 1:] = out.permute(1, 2, 0)
        return xi


class FusedRollout(pt.autograd.Function):
    """(theta, y0) -> per-path (Y_N, g(X_N), Z_sum) for a theta-independent (detached) forward process."""

    @classmethod
    def apply(cls, theta, y0, engine, call):
        call.grad_enabled = pt.is_grad_enabled()      # forward() below always runs with grad mode off
        return super().apply(theta, y0, engine, call)

    @staticmethod
    def forward(ctx, theta, y0, engine, call):
        theta_c = theta.detach().contiguous()
        y0_c = None if y0 is None else y0.detach().contiguous()
        ctx.rows_kept = engine.forward(theta_c, y0_c, call, keep_rows=ctx.needs_input_grad[0] and call.grad_enabled)
        ctx.rows_serial = engine.rows_serial
        call.X_N, call.stats = engine.X_N, engine.stats
        ctx.engine, ctx.call, ctx.has_y0 = engine, call, y0 is not None
        ctx.save_for_backward(theta_c)
        ctx.set_materialize_grads(False)      # an output the loss does not use arrives as None, not as zeros
        Y, gX, Zsum = engine.Y_N.clone(), engine.gX.clone(), engine.Zsum.clone()
        ctx.mark_non_differentiable(gX)
        return Y, gX, Zsum

    @staticmethod
    def backward(ctx, gY, ggX, gZsum):
        (theta_c,) = ctx.saved_tensors
        engine = ctx.engine
        wY = pt.zeros(engine.K_local, dtype=pt.float32, device=engine.device) if gY is None else gY.contiguous().float()
        wZ = None if gZsum is None else gZsum.contiguous().float()
        grad = pt.empty(engine.n_theta, dtype=pt.float32, device=engine.device)
        if wZ is not None:
            engine.ckpt_ok, engine.ckpt = False, None       # this loss needs the rollout with the cotangents in hand
        if ctx.rows_kept and wZ is None and engine.ckpt is not None and ctx.rows_serial == engine.rows_serial:
            engine.grad_from_rows(theta_c, wY, ctx.call, grad)      # the rows are those of THIS forward
        else:
            engine.backward_detached(theta_c, wY, wZ, ctx.call, grad)
        gy0 = wY.sum().reshape(1) if ctx.has_y0 else None
        return grad, gy0, None, None


class FusedRolloutAttached(pt.autograd.Function):
    """theta -> sum_k (Z_sum + g(X_N))_k / K_global over the local shard, attached forward process."""

    @staticmethod
    def forward(ctx, theta, engine, call):
        theta_c = theta.detach().contiguous()
        grad = pt.empty(engine.n_theta, dtype=pt.float32, device=engine.device)
        engine.attached(theta_c, call, grad)
        call.X_N, call.stats = engine.X_N, engine.stats
        ctx.save_for_backward(grad)
        return (engine.stats[2] / engine.K_global).float()

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None


class FusedRolloutAttachedGeneral(pt.autograd.Function):
    """(theta, y0) -> per-path (Y_N, g(X_N), Z_sum) for the ATTACHED adaptive forward process (detach_forward=False).
    Two-phase: forward = forward rollout kernel; backward = checkpointed adjoint kernel driven by the per-path
    cotangents (dL/dY_N, dL/dg(X_N), dL/dZsum) -- any loss of solver.py:164-192."""

    @staticmethod
    def forward(ctx, theta, y0, engine, call):
        theta_c = theta.detach().contiguous()
        y0_c = None if y0 is None else y0.detach().contiguous()
        engine.forward(theta_c, y0_c, call)
        call.X_N, call.stats = engine.X_N, engine.stats
        ctx.engine, ctx.call, ctx.has_y0, ctx.y0 = engine, call, y0 is not None, y0_c
        ctx.save_for_backward(theta_c)
        return engine.Y_N.clone(), engine.gX.clone(), engine.Zsum.clone()

    @staticmethod
    def backward(ctx, gY, ggX, gZsum):
        (theta_c,) = ctx.saved_tensors
        engine = ctx.engine
        c = lambda t: None if t is None else t.contiguous().float()
        wY, wG, wZ = c(gY), c(ggX), c(gZsum)
        if wY is None and wG is None and wZ is None:
            wY = pt.zeros(engine.K_local, dtype=pt.float32, device=engine.device)
        grad = pt.empty(engine.n_theta, dtype=pt.float32, device=engine.device)
        engine.attached(theta_c, ctx.call, grad, y0=ctx.y0, wY=wY, wZ=wZ, wG=wG)
        gy0 = (wY.sum().reshape(1) if wY is not None else pt.zeros(1, device=engine.device)) if ctx.has_y0 else None
        return grad, gy0, None, None
