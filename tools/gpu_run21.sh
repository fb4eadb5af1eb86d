#!/bin/bash
# final bench lines of HEAD for every workload + the reference arm
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for w in c2 c1 c3re c4; do echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo rc=$?; tail -c 300 gpurun_out/bench_$w.err; done
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo rc=$?
python - <<'PY'
import json
for w in ("c2","c1","c3re","c4","ref_c2"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%w).read().splitlines() if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(w, "value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), r.get("kernel_ms"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e: print(w, "ERR", e)
PY
