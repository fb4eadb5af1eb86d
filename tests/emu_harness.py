"""Drives the host-emulator build of the CUDA sources (tests/emu) through the same C ABI with numpy buffers.

Test infrastructure only: lets `-m "not gpu"` tests execute the real kernel source (index logic, shared-memory
layout, barrier placement, host-side planning) on the CPU.  The product never loads this library."""
import ctypes
import os
import subprocess

import numpy as np

from pspde import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "path-space-pde-solver_b200", "csrc")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libpspde_emu.so")
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith("_")] + [os.path.join(EMU_DIR, "simt_emul.h")]
        if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < max(os.path.getmtime(s) for s in srcs):
            os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
            cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DPSPDE_EMULATE", "-x", "c++",
                                   "-I", EMU_DIR, "-I", CSRC] + cus + ["-o", EMU_SO])
        _emu = L.bind(EMU_SO)
    return _emu


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def problem_pack(kind, d, pkw, A=None, B=None):
    """(problem_id, flags, fp32 pack) as laid out in include/pspde.h."""
    z, one = np.zeros(d, np.float32), np.ones(d, np.float32)
    flags = 0
    if kind in ("llgc", "lqgc"):
        A = -np.eye(d, dtype=np.float32) if A is None else A.astype(np.float32)
        B = np.eye(d, dtype=np.float32) if B is None else B.astype(np.float32)
        lq = kind == "lqgc"
        vecs = [np.diag(A).copy(), np.diag(B).copy(), 0.5 * one if lq else z, one if lq else z, z if lq else one, z, z]
        pack = np.concatenate(vecs)
        if np.any(A != np.diag(np.diag(A))) or np.any(B != np.diag(np.diag(B))):
            flags = L.FLAG_DENSE_AB
            pack = np.concatenate([pack, A.reshape(-1), B.reshape(-1)])
        return L.PROBLEM_OU, flags, pack.astype(np.float32)
    if kind == "dwm":
        d1, d2 = int(pkw["d_1"]), int(pkw["d_2"])
        eta = np.array([pkw["eta"]] * d1 + [1.0] * d2, np.float32)
        kap = np.array([pkw["kappa"]] * d1 + [1.0] * d2, np.float32)
        return L.PROBLEM_DW, 0, np.concatenate([z, one, z, z, z, kap, eta]).astype(np.float32)
    raise ValueError(kind)


def cfg_from_golden(g, noise=L.NOISE_INJECT, K=None, k_offset=0, seed=0, offset=0):
    d, N = g["d"], g["N"]
    pid, flags, pack = problem_pack(g["kind"], d, g.get("pkw", {}), g.get("A"), g.get("B"))
    outer = g["time_approx"] == "outer"
    dims = [d if outer else d + 1, 30, 30, d]
    cfg = L.make_cfg(K or g["K"], d, N, np.float32(g["delta_t"]), pid,
                     L.NET_DENSENET if g["net"] == "densenet" else L.NET_MLP_TANH, dims,
                     L.TIME_NONE if outer else L.TIME_FIRST, adaptive=g["adaptive"], k_offset=k_offset,
                     problem_flags=flags, noise_mode=noise, seed=seed, offset=offset,
                     xi_strides=(d * (N + 1), N + 1, 1))
    x0 = (-np.ones(d) if g["kind"] == "dwm" else np.zeros(d)).astype(np.float32)
    return cfg, pack, x0


class Runner:
    """numpy-buffer front end of the C ABI (works for the emulator; the GPU tests use torch tensors instead)."""

    def __init__(self, lib):
        self.lib = lib

    def fwd(self, cfg, theta, pack, x0, xi=None, y0=None):
        K, d = cfg.K_local, cfg.d
        ws = np.zeros(self.lib.pspde_workspace_bytes(ctypes.byref(cfg)) // 8 + 1, np.float64)
        out = dict(X=np.zeros((K, d), np.float32), Y=np.zeros(K, np.float32), gX=np.zeros(K, np.float32),
                   Zsum=np.zeros(K, np.float32), stats=np.zeros(4, np.float64))
        xi_p = None if xi is None else ctypes.c_void_p(xi.ctypes.data + 4)   # slice n+1 drives step n
        y0a = None if y0 is None else np.array([y0], np.float32)
        rc = self.lib.pspde_rollout_fwd(ctypes.byref(cfg), ptr(theta), ptr(pack), ptr(x0), ptr(y0a), xi_p,
                                        ptr(out["X"]), ptr(out["Y"]), ptr(out["gX"]), ptr(out["Zsum"]),
                                        ptr(out["stats"]), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return out

    def bwd(self, cfg, theta, pack, x0, wY, wZ=None, xi=None):
        ws = np.zeros(self.lib.pspde_workspace_bytes(ctypes.byref(cfg)) // 8 + 1, np.float64)
        grad = np.full(self.lib.pspde_theta_size(ctypes.byref(cfg)), np.nan, np.float32)
        xi_p = None if xi is None else ctypes.c_void_p(xi.ctypes.data + 4)
        wY = np.ascontiguousarray(wY, np.float32)
        wZ = None if wZ is None else np.ascontiguousarray(wZ, np.float32)
        rc = self.lib.pspde_rollout_bwd_detached(ctypes.byref(cfg), ptr(theta), ptr(pack), ptr(x0), xi_p, ptr(wY),
                                                 ptr(wZ), ptr(grad), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return grad

    def attached(self, cfg, theta, pack, x0, w, xi=None, y0=None, wY=None, wZ=None, wG=None):
        K, d = cfg.K_local, cfg.d
        ws = np.zeros(self.lib.pspde_workspace_bytes(ctypes.byref(cfg)) // 8 + 1, np.float64)
        out = dict(X=np.zeros((K, d), np.float32), Y=np.zeros(K, np.float32), gX=np.zeros(K, np.float32),
                   Zsum=np.zeros(K, np.float32), stats=np.zeros(4, np.float64),
                   grad=np.full(self.lib.pspde_theta_size(ctypes.byref(cfg)), np.nan, np.float32))
        xi_p = None if xi is None else ctypes.c_void_p(xi.ctypes.data + 4)
        y0a = None if y0 is None else np.array([y0], np.float32)
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        wY, wZ, wG = f(wY), f(wZ), f(wG)
        rc = self.lib.pspde_rollout_attached(ctypes.byref(cfg), ptr(theta), ptr(pack), ptr(x0), ptr(y0a), xi_p,
                                             ctypes.c_float(w), ptr(wY), ptr(wZ), ptr(wG), ptr(out["X"]),
                                             ptr(out["Y"]), ptr(out["gX"]), ptr(out["Zsum"]), ptr(out["stats"]),
                                             ptr(out["grad"]), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return out


def diffusion_cfg(K, d, N, dt, arch, noise=L.NOISE_INJECT, k_offset=0, seed=0, offset=0, problem_id=L.PROBLEM_HEAT):
    dims = [d + 1] + [int(a) for a in arch] + [1]
    return L.make_cfg(K, d, N, np.float32(dt), problem_id, L.NET_DENSENET, dims, L.TIME_LAST, adaptive=False,
                      k_offset=k_offset, noise_mode=noise, seed=seed, offset=offset, xi_strides=(d, 1, K * d), n_sets=0)


def heat_pack(d):
    z = np.zeros(d, np.float32)
    return np.concatenate([z, np.full(d, np.sqrt(np.float32(2.0)), np.float32), z, z, z, z, z]).astype(np.float32)


class DiffusionRunner:
    """numpy front end of pspde_diffusion_* (emulator build)."""

    def __init__(self, lib):
        self.lib = lib

    def _ws(self, cfg, T):
        n = self.lib.pspde_diffusion_workspace_bytes(ctypes.byref(cfg), ctypes.c_float(T))
        assert n > 0, self.lib.pspde_last_error()
        return np.zeros(n // 8 + 1, np.float64)

    def fwd(self, cfg, T, theta, pack, X0, t0, xis=None):
        K, d = cfg.K_local, cfg.d
        ws = self._ws(cfg, T)
        o = dict(V0=np.zeros(K, np.float32), VE=np.zeros(K, np.float32), Y=np.zeros(K, np.float32),
                 X=np.zeros((K, d), np.float32), t=np.zeros(K, np.float32), stats=np.zeros(4, np.float64))
        rc = self.lib.pspde_diffusion_fwd(ctypes.byref(cfg), ctypes.c_float(T), ptr(theta), ptr(pack), ptr(X0), ptr(t0),
                                          ptr(xis), ptr(o["V0"]), ptr(o["VE"]), ptr(o["Y"]), ptr(o["X"]), ptr(o["t"]),
                                          ptr(o["stats"]), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return o

    def bwd(self, cfg, T, theta, pack, X0, t0, xis, c0, cE, cD):
        ws = self._ws(cfg, T)
        grad = np.full(self.lib.pspde_theta_size(ctypes.byref(cfg)), np.nan, np.float32)
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        c0, cE, cD = f(c0), f(cE), f(cD)
        rc = self.lib.pspde_diffusion_bwd(ctypes.byref(cfg), ctypes.c_float(T), ptr(theta), ptr(pack), ptr(X0), ptr(t0),
                                          ptr(xis), ptr(c0), ptr(cE), ptr(cD), ptr(grad), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return grad

    def iteration(self, g, theta, K_boundary, alpha=(1.0, 1.0, 1.0), T=1.0, kind="heat"):
        """loss, grad, K_count, per-path outputs of one GeneralSolver iteration (solver.py:1062-1064, :1076-1163);
        kind 'heat' (h == 0, terminal |x|^2) or 'allencahn' (h = y - y^3, terminal 1 / (2 + 2/5 |x|^2))."""
        pid = L.PROBLEM_HEAT if kind == "heat" else L.PROBLEM_ALLEN_CAHN
        f_term = (lambda x: (x ** 2).sum(1)) if kind == "heat" else (lambda x: 1 / (2 + 2 / 5 * (x ** 2).sum(1)))
        K, d, N = int(g["K"]), int(g["d"]), int(g["N"])
        X0 = np.ascontiguousarray(g["X0"], np.float32)
        t0 = np.ascontiguousarray(g["t0"], np.float32).reshape(-1)
        xis = np.ascontiguousarray(g["xis"], np.float32)
        pack = heat_pack(d)
        cfg = diffusion_cfg(K, d, N, g["delta_t"], g["arch"], problem_id=pid)
        f = self.fwd(cfg, T, theta, pack, X0, t0, xis)
        r = f["VE"].astype(np.float64) - f["Y"]
        loss = alpha[0] * np.mean(r ** 2)
        w = alpha[0] * 2 * r / K
        grad = self.bwd(cfg, T, theta, pack, X0, t0, xis, -w, w, -w).astype(np.float64)
        # terminal condition on the first K_boundary samples (N = 0 call at t = T)
        Kb = K_boundary
        cfgb = diffusion_cfg(Kb, d, 0, g["delta_t"], g["arch"], problem_id=pid)
        Xb, tb = np.ascontiguousarray(X0[:Kb]), np.full(Kb, T, np.float32)
        fb = self.fwd(cfgb, T, theta, pack, Xb, tb)
        rT = fb["V0"].astype(np.float64) - f_term(Xb.astype(np.float64))
        loss += alpha[1] * np.mean(rT ** 2)
        grad += self.bwd(cfgb, T, theta, pack, Xb, tb, None, alpha[1] * 2 * rT / Kb, None, None)
        return dict(loss=loss, grad=grad, K_count=int(f["stats"][1]), **f)


# ---------------------------------------------------------------------------------------------- elliptic (row f4)
ELLIPTIC_H = {"expsphere": L.H_EXP_LINEAR, "expball": L.H_EXP_NONLINEAR, "expball_sin": L.H_EXP_NONLINEAR_SIN,
              "helmholtz": L.H_HELMHOLTZ}


def elliptic_spec(kind, alpha=1.0):
    """pspde_elliptic of the reference problems (problems.py:962-1064 unit ball; :1614-1654 square [-1, 1]^2)."""
    if kind == "helmholtz":
        return L.make_elliptic(L.DOMAIN_BOX, x_l=-1.0, x_r=1.0, one_boundary=False, h_id=L.H_HELMHOLTZ,
                               h_param=(1.0, 1.0, 4.0))
    if kind == "committor":
        return L.make_elliptic(L.DOMAIN_ANNULUS, radius=2.0, radius_in=1.0, h_id=L.H_COMMITTOR, h_param=(1.0, 2.0, 0.0))
    return L.make_elliptic(L.DOMAIN_SPHERE, radius=1.0, h_id=ELLIPTIC_H[kind], h_param=(alpha, 0.0, 0.0))


def elliptic_cfg(K, d, N, dt, arch, noise=L.NOISE_INJECT, k_offset=0, seed=0, offset=0):
    dims = [d] + [int(a) for a in arch] + [1]
    return L.make_cfg(K, d, N, np.float32(dt), L.PROBLEM_HEAT, L.NET_DENSENET, dims, L.TIME_NONE, adaptive=False,
                      k_offset=k_offset, noise_mode=noise, seed=seed, offset=offset, xi_strides=(d, 1, K * d), n_sets=1)


class EllipticRunner:
    """numpy front end of pspde_elliptic_* (emulator build)."""

    def __init__(self, lib):
        self.lib = lib

    def _ws(self, cfg, ell):
        n = self.lib.pspde_elliptic_workspace_bytes(ctypes.byref(cfg), ctypes.byref(ell))
        assert n > 0, self.lib.pspde_last_error()
        return np.zeros(n // 8 + 1, np.float64)

    def fwd(self, cfg, ell, theta, pack, X0, xis=None):
        K, d = cfg.K_local, cfg.d
        ws = self._ws(cfg, ell)
        o = dict(V0=np.zeros(K, np.float32), VE=np.zeros(K, np.float32), Y=np.zeros(K, np.float32),
                 X=np.zeros((K, d), np.float32), VL2=np.zeros(K, np.float32), stats=np.zeros(4, np.float64))
        rc = self.lib.pspde_elliptic_fwd(ctypes.byref(cfg), ctypes.byref(ell), ptr(theta), ptr(pack), ptr(X0), ptr(xis),
                                         ptr(o["V0"]), ptr(o["VE"]), ptr(o["Y"]), ptr(o["X"]), ptr(o["VL2"]),
                                         ptr(o["stats"]), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return o

    def bwd(self, cfg, ell, theta, pack, X0, xis, c0, cE, cD):
        ws = self._ws(cfg, ell)
        grad = np.full(self.lib.pspde_theta_size(ctypes.byref(cfg)), np.nan, np.float32)
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        c0, cE, cD = f(c0), f(cE), f(cD)
        rc = self.lib.pspde_elliptic_bwd(ctypes.byref(cfg), ctypes.byref(ell), ptr(theta), ptr(pack), ptr(X0), ptr(xis),
                                         ptr(c0), ptr(cE), ptr(cD), ptr(grad), ptr(ws), ws.nbytes, None)
        L.check(self.lib, rc)
        return grad

    def iteration(self, g, theta, g_fun, alpha=(1.0, 1.0)):
        """loss, grad, K_count, per-path outputs of one EllipticSolver iteration (solver.py:646-670, :687-790)."""
        K, d, N = int(g["K"]), int(g["d"]), int(g["N"])
        X0 = np.ascontiguousarray(g["X0"], np.float32)
        Xb = np.ascontiguousarray(g["Xb"], np.float32)
        xis = np.ascontiguousarray(g["xis"], np.float32)
        pack = heat_pack(d)
        if str(g["kind"]) == "committor":              # sigma = I
            pack[d:2 * d] = 1.0
        ell = elliptic_spec(str(g["kind"]))
        K = X0.shape[0]                                # 'two_spheres': the start points outside the annulus were dropped
        cfg = elliptic_cfg(K, d, N, g["delta_t"], g["arch"])
        f = self.fwd(cfg, ell, theta, pack, X0, xis)
        r = f["VE"].astype(np.float64) - f["Y"]
        loss = alpha[0] * np.mean(r ** 2)
        w = alpha[0] * 2 * r / K
        grad = self.bwd(cfg, ell, theta, pack, X0, xis, -w, w, -w).astype(np.float64)
        Kb = Xb.shape[0]                                   # Dirichlet term: N = 0 call on the boundary samples
        cfgb = elliptic_cfg(Kb, d, 0, g["delta_t"], g["arch"])
        fb = self.fwd(cfgb, ell, theta, pack, Xb)
        rb = fb["V0"].astype(np.float64) - g_fun(Xb.astype(np.float64))
        loss += alpha[1] * np.mean(rb ** 2)
        grad += self.bwd(cfgb, ell, theta, pack, Xb, None, None, alpha[1] * 2 * rb / Kb, None)
        return dict(loss=loss, grad=grad, K_count=int(f["stats"][1]), **f)
