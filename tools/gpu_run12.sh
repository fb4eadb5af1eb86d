#!/bin/bash
# consolidation run on one B200: tests, C2/C3 bench (+ reference arm), launch list, full ncu captures of the backward kernels, phase profiles
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err
echo "== bench reference arm c2"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo rc=$?
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo rc=$?
python - <<'PY'
import json
for w in ("n1","c3","ref_c2"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%w).read().splitlines() if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(w, "value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s ckpt %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), r.get("kernel_ms"), r.get("frac"), r.get("bwd_checkpoint")))
    except Exception as e: print(w, "ERR", e)
PY
echo "== phase profiles"; timeout 300 python tools/phase_profile_tc.py c2 2>&1 | tail -9 | tee gpurun_out/phase_profile_tc_fwd_c2.txt; timeout 300 python tools/bench_grad.py 100 2>&1 | tail -10 | tee gpurun_out/phase_profile_grad_c2.txt
echo "== ncu launches c2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1b_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo rc=$?
echo "== ncu full c2 (ckpt rollout + gradient kernel)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"^rollout_tc_fwd_kernel|^grad_tc_kernel" -s 9 -c 2 -o gpurun_out/prof_r1b_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_f2.log
