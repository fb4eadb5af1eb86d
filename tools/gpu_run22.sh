#!/bin/bash
# RED flush of the FMA weight gradient: all GPU tests, C1 ('outer') and C3re / C4-independent spot benches
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in c1 c3re; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; done
python - <<'PY'
import json
for w in ("c1","c3re"):
    try:
        d=json.loads([l for l in open("gpurun_out/bench_%s.json"%w).read().splitlines() if l.startswith("{")][-1])
        r=d.get("roofline") or {}
        print(w, "value %.4g ms/step %.2f e2e %.4g kernels %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_ms"), (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e: print(w, "ERR", e)
PY
