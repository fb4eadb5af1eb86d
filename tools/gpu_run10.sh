#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== gradtc"; timeout 300 python tools/check_grad_kernels.py 2>&1 | tail -6
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n1.json").read().splitlines() if l.startswith("{")][-1])
r=d.get("roofline") or {}
print("value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s loss %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), r.get("kernel_ms"), r.get("frac"), d.get("final_loss")))
PY
echo "== ncu launches c2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_gradtc_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo rc=$?
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_gradtc_c2.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4][:60]; val=float(r[-1].replace(",",""))
    if "pspde" not in name: continue
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=val
for k,(n,t) in agg.items(): print("%-62s n=%3d total=%.3f ms"%(k,n,t/1e6))
PY
