#!/bin/bash
# final round-1 evidence for the committed kernels: launch list + full ncu capture of the gradient kernel and the checkpoint rollout in the bench pipeline
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== bench c2 (+cpu baseline)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo rc=$?
python - <<'PY'
import json
for w in ("n1","c3"):
    d=json.loads([l for l in open("gpurun_out/bench_%s.json"%w).read().splitlines() if l.startswith("{")][-1])
    r=d.get("roofline") or {}
    print(w, "value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), r.get("kernel_ms"), r.get("frac")))
PY
echo "== grad microbench"; timeout 300 python tools/bench_grad.py 100 2>&1 | tail -11 | tee gpurun_out/phase_profile_grad_c2.txt
echo "== ncu launches c2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1c_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo rc=$?
echo "== ncu full c2 (gradient kernel + checkpoint rollout, in the bench pipeline)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"^grad_tc_kernel|^rollout_tc_fwd_kernel" -s 8 -c 3 -o gpurun_out/prof_r1c_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_f2.log
