#!/usr/bin/env python
"""bench.py -- path-steps/sec per training iteration of the fused path-space rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c3re|c4|c5]

Workload (N=1 default): BASELINE.json configs[1] -- Ornstein-Uhlenbeck HJB with linear costs, LLGC(d=100,
off_diag=0, T=1), DenseNet(101 -> 30 -> 30 -> 100) 'inner', K = 2^16 trajectories per GPU, delta_t = 0.01
(N = 100 steps), log-variance loss, detach_forward=True.  One "step" = one full training iteration: in-kernel
Philox noise, forward rollout, loss statistics, all-reduce of the statistics (N > 1), backward rollout
(recompute), all-reduce of the gradient (N > 1), Adam.  Weak scaling: K per GPU is fixed, K_global = N * 2^16.

Arms
  ours        the CUDA path through pspde.Solver (fails without a GPU: there is no CPU fallback)
  reference   the reference's CPU algorithm (oracle/ref_port.py: PyTorch eager + autograd + CPU randn, the exact
              procedure of solver.py:420-531) on the host cores, on a bounded sample of the same workload
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "path-space-pde-solver_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch as pt  # noqa: E402

METRIC = "path-steps/sec per train iter"
UNIT = "path-steps/s"

WORKLOADS = {
    # name: (problem kind, ctor kwargs, net, time_approx, K per GPU, delta_t, loss, detach, lr)
    "c2": dict(kind="llgc", pkw=dict(d=100, off_diag=0, T=1, seed=42), net="densenet", ta="inner", K=1 << 16,
               dt=0.01, loss="log-variance", detach=True, lr=1e-3,
               desc="C2: LLGC d=100 off_diag=0 T=1, DenseNet[101,30,30,100] inner, K=2^16/GPU, N=100, log-variance"),
    "c5": dict(kind="llgc", pkw=dict(d=100, off_diag=0, T=1, seed=42), net="densenet", ta="inner", K=1 << 20,
               dt=0.005, loss="log-variance", detach=True, lr=1e-3,
               desc="C5: LLGC d=100, DenseNet inner, K=2^20/GPU, N=200, log-variance"),
    "c1": dict(kind="lqgc", pkw=dict(d=10), net="densenet", ta="outer", K=200, dt=0.05, loss="log-variance",
               detach=True, lr=1e-3, desc="C1: LQGC d=10, 100 x DenseNet[10,30,30,10] outer, K=200, N=100"),
    "c3": dict(kind="dwm", pkw=dict(d=50, d_1=15, d_2=35, T=1, eta=3, kappa=5), net="mlp", ta="inner", K=1 << 18,
               dt=0.005, loss="log-variance", detach=True, lr=0.05,
               desc="C3: DoubleWell_multidim d=50, MySequential, K=2^18/GPU, N=200, log-variance"),
    "c3re": dict(kind="dwm", pkw=dict(d=50, d_1=15, d_2=35, T=1, eta=3, kappa=5), net="mlp", ta="inner", K=1 << 18,
                 dt=0.005, loss="relative_entropy", detach=False, lr=0.05,
                 desc="C3: DoubleWell_multidim d=50, MySequential, K=2^18/GPU, N=200, relative entropy (attached)"),
    "c4": dict(kind="heat", pkw=dict(d=50, T=1), net="densenet", arch=[256, 256], K=1 << 18, K_boundary=50, N=25,
               dt=1e-3, loss="diffusion", lr=1e-3,
               desc="C4: HeatEquation d=50, GeneralSolver diffusion loss, DenseNet[51,256,256,1], K=2^18/GPU, N=25"),
}


def net_macs(dims, dense):
    """forward MACs M, hidden-cotangent MACs M_delta."""
    L = len(dims) - 1
    if dense:
        M = sum(sum(dims[:i + 1]) * dims[i + 1] for i in range(L))
        Md = sum(sum(dims[1:i + 1]) * dims[i + 1] for i in range(1, L))
    else:
        M = sum(dims[i] * dims[i + 1] for i in range(L))
        Md = sum(dims[i] * dims[i + 1] for i in range(1, L))
    return M, Md


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, wl):
    """The reference's CPU procedure (torch eager, CPU randn, autograd, Adam) on a bounded sample."""
    from oracle import ref_port as orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pt.set_num_threads(cores)
    d = wl["pkw"]["d"]
    kind = wl["kind"]
    pkw = {k: v for k, v in wl["pkw"].items() if k != "d"}
    prob = orc.make_problem(kind, d, **pkw)
    if kind == "heat":
        N, K_cpu = wl["N"], min(wl["K"], 2048)
        params = orc.densenet_init(d + 1, 1, wl["arch"], seed=42)
        times = []
        orc.diffusion_train_loop(prob, params, K_cpu, wl["K_boundary"], N, wl["dt"], args.warmup + args.steps, wl["lr"],
                                 seed=42, times=times)
        return _print_reference(args, wl, times, K_cpu, N, cores)
    N = int(np.floor(prob.T / wl["dt"]))
    K_cpu = min(wl["K"], 4096)      # the reference pre-draws xi = randn(K, d, N+1) on the host: ~0.5 MB per path
    if wl["net"] == "mlp":
        params, net = orc.mlp_init(d + 1, d, seed=123), "mlp_tanh"
    elif wl["ta"] == "outer":
        params, net = [orc.densenet_init(d, d, seed=42) for _ in range(N)], "densenet"
    else:
        params, net = orc.densenet_init(d + 1, d, seed=42), "densenet"
    times = []
    orc.hjb_train_loop(prob, net, params, K_cpu, wl["dt"], args.warmup + args.steps, wl["lr"], wl["loss"], wl["ta"],
                       True, wl["detach"], seed=42, times=times)
    _print_reference(args, wl, times, K_cpu, N, cores)


def _print_reference(args, wl, times, K_cpu, N, cores):
    t = times[args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    value = K_cpu * N / (ms * 1e-3)
    sample = "K=%d of %d trajectories per step, N=%d, %d timed iterations" % (K_cpu, wl["K"], N, len(t))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "device": "cpu", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- our arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--id=%d" % index, "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(power))
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except OSError:
        return {}


def fma_peak_tflops(lib, dev):
    """FP32 FMA throughput of this GPU, measured live with the library's probe kernel (best of 5)."""
    sink = pt.zeros(4, device=dev)
    stream = ctypes.c_void_p(pt.cuda.current_stream(dev).cuda_stream)
    best = 0.0
    for _ in range(6):
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.pspde_fma_probe(20000, ctypes.c_void_p(sink.data_ptr()), stream)
        e1.record()
        e1.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def fwd_roofline(flops_fwd, tf_ms, fma_peak, eng):
    """Forward kernel: on the tensor cores (tcgen05 kind::tf32, 3 passes per product for FP32 equivalence) when the
    network is in the tensor-core kernel's shape class, else FP32 FMA.  The tensor bound is the measured bf16 GEMM
    peak / 2 (TF32 runs at half the bf16 rate) / 3 (passes): FP32-equivalent TFLOP/s."""
    ach = flops_fwd / (tf_ms * 1e-3) / 1e12
    hid_max = 32 if eng.net_id == 0 else 31          # MySequential keeps a bias column per hidden segment
    tc_path = os.environ.get("PSPDE_FWD_PATH", "") != "simt" and len(eng.dims) == 4 \
        and max(eng.dims[1:3]) <= hid_max and eng.time_mode == 0 and not (eng.flags & 1)
    out = {"achieved": ach, "frac_of_fp32_fma_peak": ach / fma_peak}
    if tc_path:
        bf16 = measured_peaks().get("bf16_tflops")
        src = "MEASURED_PEAKS.json bf16_tflops (burst) / 2 / 3"
        if not bf16:
            bf16, src = 1590.0, "fallback 1.59 PFLOP/s bf16 / 2 / 3"
        out.update(kernel="rollout_tc_fwd_kernel (tcgen05.mma kind::tf32, 3xTF32, A from tensor memory)", bound="tensor",
                   peak=bf16 / 6.0, frac=ach / (bf16 / 6.0), peak_source=src)
    else:
        out.update(kernel="rollout_kernel<FWD>", bound="fp32_fma", peak=fma_peak, frac=ach / fma_peak)
    return out


def build_solver(wl, K_global, dev):
    import pspde
    pkw = dict(wl["pkw"])
    d = pkw["d"]
    if wl["kind"] == "llgc":
        prob = pspde.LLGC(device=dev, **pkw)
    elif wl["kind"] == "lqgc":
        prob = pspde.LQGC(device=dev, **pkw)
    else:
        prob = pspde.DoubleWell_multidim(device=dev, **pkw)
    S = pspde.Solver("bench", prob, lr=wl["lr"], L=1, K=K_global, delta_t=wl["dt"], loss_method=wl["loss"],
                     time_approx=wl["ta"], detach_forward=wl["detach"], early_stopping_time=None,
                     u_l2_error_flag=False, verbose=False, seed=42, noise="philox", device=dev)
    if wl["net"] == "densenet" and wl["ta"] == "inner":
        S.z_n = pspde.DenseNet(d_in=d + 1, d_out=d, lr=wl["lr"], seed=42)
        S.update_Phis()
    return S


def run_ours(args, wl):
    import torch.distributed as td
    from pspde import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not pt.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    pt.cuda.set_device(local)
    dev = pt.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    K_global = wl["K"] * world
    S = build_solver(wl, K_global, dev)
    N, d = S.N, S.d
    eng = S._get_engine()
    flush = pt.empty(256 << 20, dtype=pt.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            td.barrier()
        pt.cuda.synchronize(dev)

    # ---- kernel-level timing of the two rollout kernels (for the roofline), on the launching stream
    def time_kernels(reps):
        tf, tb = [], []
        theta = S._theta.detach()
        call = type("C", (), {})()
        from pspde.fused import Call
        wY = pt.randn(eng.K_local, device=dev) / K_global
        grad = pt.empty(eng.n_theta, device=dev)
        for i in range(reps):
            c = Call(offset=1000 + i)
            flush.fill_(i & 1)
            e = [pt.cuda.Event(enable_timing=True) for _ in range(4)]
            single = eng.ckpt is not None          # the training step keeps the forward's operand rows (single rollout)
            e[0].record(); eng.forward(theta, None, c, keep_rows=single); e[1].record()
            flush.fill_(1 - (i & 1))
            e[2].record()
            if single:
                eng.grad_from_rows(theta, wY, c, grad)
            else:
                eng.backward_detached(theta, wY, None, c, grad)
            e[3].record()
            pt.cuda.synchronize(dev)
            tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
        return tf, tb

    for l in range(args.warmup):
        S.train_step(l)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.pspde_launch_count()
    # ---- timed region A: device-timed steps (per-step CUDA event pairs, L2 flushed between steps)
    step_ms = []
    for l in range(args.steps):
        flush.fill_(l & 1)
        barrier()
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record()
        S.train_step(args.warmup + l)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = lib.pspde_launch_count() - launches0
    t = pt.tensor([sum(step_ms)], dtype=pt.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = K_global * N * args.steps / (total_ms * 1e-3)

    # ---- timed region B: end to end through the public API with host inputs (pinned H2D each step, D2H of the loss)
    pack_h = eng.pack.detach().cpu().pin_memory()
    x0_h = eng.x0.detach().cpu().pin_memory()
    barrier()
    w0 = time.perf_counter()
    for l in range(args.steps):
        eng.pack.copy_(pack_h, non_blocking=True)
        eng.x0.copy_(x0_h, non_blocking=True)
        S.train_step(args.warmup + args.steps + l)          # ends with loss.item(): D2H of the step's result
    barrier()
    w1 = time.perf_counter()
    t = pt.tensor([w1 - w0], dtype=pt.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_value = K_global * N * args.steps / float(t.item())
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (backward rollout), FP32 FMA bound
    M, Md = net_macs(eng.dims, eng.net_id == _lib.NET_DENSENET)
    flops_fwd = 2.0 * M * eng.K_local * N                  # forward kernel: one network evaluation per path-step
    flops_bwd = 2.0 * (M + Md) * eng.K_local * N           # backward kernel: weight gradient + hidden cotangents
    roof = None                                            # (the recomputed forward inside it is NOT counted)
    if wl["detach"]:
        tf, tb = time_kernels(5)
        tf, tb = statistics.median(tf), statistics.median(tb)
        peak = fma_peak_tflops(lib, dev)
        nominal = 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12 if clocks else None
        ach = flops_bwd / (tb * 1e-3) / 1e12
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                prof = json.load(fh)
        except OSError:
            pass
        hid_max = 32 if eng.net_id == 0 else 31
        ckpt_path = os.environ.get("PSPDE_BWD_PATH", "") != "simt" and len(eng.dims) == 4 \
            and max(eng.dims[1:3]) <= hid_max and eng.time_mode == 0 and not (eng.flags & 1)
        single = eng.ckpt is not None
        if ckpt_path and single:
            s0 = (d + 2 + 7) // 8 * 8
            ckpt_bytes = 2.0 * eng.K_local * N * (2 * (s0 // 4) + 16) * 16          # operand rows written once, read once
            kernel = ("single-rollout step: the training forward (rollout_tc_fwd_kernel<CKPT>, timed as 'fwd') keeps the "
                      "operand rows [a0|h1|h2|sqrt(dt) xi] of all tiles in HBM; backward = grad_tc_kernel over those rows "
                      "(dL/dY_N applied at load, hidden cotangents in FP32 FMA, weight gradient on tcgen05 kind::tf32 3xTF32 "
                      "with K = samples, accumulators resident in tensor memory) + reduce; tiles the buffer does not hold "
                      "(rows_kept_fraction < 1) are recomputed by the wave-checkpointed backward inside 'bwd'")
        elif ckpt_path:
            s0 = (d + 2 + 7) // 8 * 8
            ckpt_bytes = 2.0 * eng.K_local * N * (2 * (s0 // 4) + 16) * 16          # operand rows written once, read once
            kernel = ("checkpointed detached backward = rollout_tc_fwd_kernel<CKPT> (tensor-core rollout that leaves the "
                      "operand rows [a0|h1|h2|zeta] of one wave of tiles in the workspace) + grad_tc_kernel (hidden "
                      "cotangents in FP32 FMA, weight gradient on tcgen05 kind::tf32 3xTF32 with K = samples, accumulators "
                      "resident in tensor memory); timed together")
        else:
            ckpt_bytes, kernel = 0.0, "rollout_kernel<BWD> (detached backward, FP32-FMA recompute)"
        kept = (eng.ckpt.numel() / eng._ckpt_need) if (ckpt_path and single) else 0.0
        roof = {"bound": "fp32_fma", "kernel": kernel,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": "fp32 FMA probe measured live in this run (MEASURED_PEAKS.json has no fp32 figure); "
                               "nominal 148 SM x 128 lanes x 2 x clocks.max.sm = %.1f TFLOP/s" % (nominal or 0),
                "traffic": prof.get(args.workload if (single or not ckpt_path) else args.workload + "_two_rollout_step",
                                    {}).get("bwd_dram_bytes_per_launch"),
                "kernel_ms": {"fwd": tf, "bwd": tb},
                "single_rollout": bool(ckpt_path and single), "rows_kept_fraction": kept,
                "bwd_checkpoint": {"bytes_per_step": ckpt_bytes,
                                   # bwd reads every row once and writes those the forward did not keep
                                   "gbs_over_bwd": ckpt_bytes / 2 * (2.0 - kept) / (tb * 1e-3) / 1e9,
                                   "gbs_over_fwd": ckpt_bytes / 2 * kept / (tf * 1e-3) / 1e9,
                                   "hbm_peak_gbs_measured": measured_peaks().get("hbm_gbs")} if ckpt_path else None,
                "fwd": fwd_roofline(flops_fwd, tf, peak, eng),
                "step": {"algorithmic_flops_per_path_step": 2.0 * (2 * M + Md),
                         "achieved": (flops_fwd + flops_bwd) / ((tf + tb) * 1e-3) / 1e12,
                         "frac": (flops_fwd + flops_bwd) / ((tf + tb) * 1e-3) / 1e12 / peak},
                "hbm_peak_gbs_measured": measured_peaks().get("hbm_gbs")}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_port as orc
        cores = os.cpu_count() or 1
        pt.set_num_threads(cores)
        pkw = {k: v for k, v in wl["pkw"].items() if k != "d"}
        prob = orc.make_problem(wl["kind"], d, **pkw)
        K_cpu = min(wl["K"], 4096)
        if wl["net"] == "mlp":
            params, net = orc.mlp_init(d + 1, d, seed=123), "mlp_tanh"
        elif wl["ta"] == "outer":
            params, net = [orc.densenet_init(d, d, seed=42) for _ in range(N)], "densenet"
        else:
            params, net = orc.densenet_init(d + 1, d, seed=42), "densenet"
        times = []
        orc.hjb_train_loop(prob, net, params, K_cpu, wl["dt"], 4, wl["lr"], wl["loss"], wl["ta"], True, wl["detach"],
                           seed=42, times=times)
        tt = times[1:]
        cpu = {"value": K_cpu * N / (sum(tt) / len(tt)), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "K=%d of %d trajectories, N=%d, 3 timed iterations after 1 warm-up (torch CPU eager + "
                         "autograd + CPU randn, the reference's procedure)" % (K_cpu, wl["K"], N)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "K_global": K_global, "K_per_gpu": wl["K"], "N": N, "d": d,
                       "noise": "in-kernel Philox4x32-10", "parallelism": "trajectory-sharded dp%d" % world,
                       "l2": "256 MiB buffer written between timed steps (outside the per-step CUDA-event pairs); "
                             "the kernels' working set is on-chip"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pack_h.numel() * 4 + x0_h.numel() * 4),
                    "d2h_bytes_per_step": 8,
                    "how": "pspde.Solver.train_step through the C ABI, wall clock, problem functor pack + x0 copied "
                           "from pinned host memory and the loss read back every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "final_loss": S.loss_log[-1]}
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def run_ours_c4(args, wl):
    """BASELINE config 4: one GeneralSolver iteration (device-side sampling, forward rollout with the directional
    derivative of V, loss, backward rollout, terminal-condition term, Adam) per step."""
    import torch.distributed as td
    import pspde
    from pspde import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not pt.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    pt.cuda.set_device(local)
    dev = pt.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    d, N = wl["pkw"]["d"], wl["N"]
    K_global = wl["K"] * world
    prob = pspde.HeatEquation(device=dev, **wl["pkw"])
    G = pspde.GeneralSolver(prob, "bench", seed=42, delta_t=wl["dt"], N=N, lr=wl["lr"], L=1, K=K_global,
                            K_boundary=wl["K_boundary"], verbose=False, device=dev)
    G.V = pspde.DenseNet(d_in=d + 1, d_out=1, lr=wl["lr"], arch=wl["arch"], seed=42)
    eng = G._get_engine()
    flush = pt.empty(256 << 20, dtype=pt.uint8, device=dev)

    def barrier():
        if world > 1:
            td.barrier()
        pt.cuda.synchronize(dev)

    for l in range(args.warmup):
        G.train_step(l)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.pspde_launch_count()
    step_ms = []
    for l in range(args.steps):
        flush.fill_(l & 1)
        barrier()
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record()
        G.train_step(args.warmup + l)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    launches = lib.pspde_launch_count() - launches0
    t = pt.tensor([sum(step_ms)], dtype=pt.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    total_ms = float(t.item())
    value = K_global * N * args.steps / (total_ms * 1e-3)
    pack_h = eng.pack.detach().cpu().pin_memory()
    barrier()
    w0 = time.perf_counter()
    for l in range(args.steps):
        eng.pack.copy_(pack_h, non_blocking=True)
        G.train_step(args.warmup + args.steps + l)
    barrier()
    w1 = time.perf_counter()
    t = pt.tensor([w1 - w0], dtype=pt.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_value = K_global * N * args.steps / float(t.item())
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    # kernel-level timing + roofline (FP32 FMA bound; value and tangent rows: 2 network rows per path-step)
    M, Md = net_macs(eng.dims, True)
    theta = G._theta.detach()
    X0, t0 = eng.sample(1.0, 12345)
    w = pt.randn(eng.K_local, device=dev) / K_global
    grad = pt.empty(eng.n_theta, device=dev)
    tf, tb = [], []
    for i in range(4):
        flush.fill_(i & 1)
        e = [pt.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); eng.forward(theta, X0, t0, None, 12345); e[1].record()
        flush.fill_(1 - (i & 1))
        e[2].record(); eng.backward(theta, X0, t0, None, 12345, -w, w, -w, grad); e[3].record()
        pt.cuda.synchronize(dev)
        if i:
            tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
    tf, tb = statistics.median(tf), statistics.median(tb)
    peak = fma_peak_tflops(lib, dev)
    ps = eng.K_local * N
    flops_fwd, flops_bwd = 4.0 * M * ps, 4.0 * (M + Md) * ps
    ach = flops_bwd / (tb * 1e-3) / 1e12
    roof = {"bound": "fp32_fma", "kernel": "diffusion_kernel<BWD> (reverse of the value/tangent pair + weight gradient, recompute)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": "fp32 FMA probe measured live in this run (MEASURED_PEAKS.json has no fp32 figure)",
            "traffic": None, "kernel_ms": {"fwd": tf, "bwd": tb},
            "fwd": {"achieved": flops_fwd / (tf * 1e-3) / 1e12, "frac": flops_fwd / (tf * 1e-3) / 1e12 / peak},
            "step": {"algorithmic_flops_per_path_step": 4.0 * (2 * M + Md),
                     "achieved": (flops_fwd + flops_bwd) / ((tf + tb) * 1e-3) / 1e12,
                     "frac": (flops_fwd + flops_bwd) / ((tf + tb) * 1e-3) / 1e12 / peak},
            "hbm_peak_gbs_measured": measured_peaks().get("hbm_gbs")}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_port as orc
        cores = os.cpu_count() or 1
        pt.set_num_threads(cores)
        K_cpu = 1024
        params = orc.densenet_init(d + 1, 1, wl["arch"], seed=42)
        times = []
        orc.diffusion_train_loop(orc.make_problem("heat", d, T=1), params, K_cpu, wl["K_boundary"], N, wl["dt"], 3,
                                 wl["lr"], seed=42, times=times)
        tt = times[1:]
        cpu = {"value": K_cpu * N / (sum(tt) / len(tt)), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "K=%d of %d points, N=%d, 2 timed iterations after 1 warm-up (torch CPU eager, autograd with "
                         "create_graph per step: the reference's procedure)" % (K_cpu, wl["K"], N)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "K_global": K_global, "K_per_gpu": wl["K"], "N": N, "d": d,
                       "noise": "device-side Philox4x32-10 (initial points and increments)",
                       "parallelism": "trajectory-sharded dp%d" % world,
                       "l2": "256 MiB buffer written between timed steps; weights (371 KB) are re-read from L2 every tile-step by design"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pack_h.numel() * 4), "d2h_bytes_per_step": 32,
                    "how": "pspde.GeneralSolver.train_step through the C ABI, wall clock, problem functor pack copied from "
                           "pinned host memory and (loss, K_count) read back every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "final_loss": G.loss_log[-1]}
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl["kind"] == "heat":
        run_ours_c4(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
