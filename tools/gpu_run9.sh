#!/bin/bash
# checkpointed backward bring-up on one B200
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== pytest ckpt"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "checkpointed" > gpurun_out/pytest_ckpt.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_ckpt.log
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n1.json").read().splitlines() if l.startswith("{")][-1])
r=d.get("roofline") or {}
print("value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), r.get("kernel_ms"), r.get("frac")))
PY
echo "== bench c2 simt bwd"; PSPDE_BWD_PATH=simt timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('simt: ms/step %.2f kernels %s' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
echo "== ncu launches c2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ckpt_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo rc=$?
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_ckpt_c2.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4][:60]; val=float(r[-1].replace(",",""))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=val
for k,(n,t) in agg.items(): print("%-62s n=%3d total=%.3f ms"%(k,n,t/1e6))
PY
