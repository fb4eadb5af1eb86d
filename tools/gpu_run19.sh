#!/bin/bash
# partial forward checkpoint: tests, then C5 (K = 2^20: buffer holds part of the tiles), C3 and C2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_rollout or checkpointed" 2>&1 | tail -8
for w in c5 c3 c2; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${w}_rows.json 2> gpurun_out/bench_${w}_rows.err || tail -5 gpurun_out/bench_${w}_rows.err
done
python - <<'PY'
import json
for n in ("c5", "c3", "c2"):
    try:
        j = json.loads(open("gpurun_out/bench_%s_rows.json" % n).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(n, "%.4g" % j["value"], "%.2f ms" % j["ms_per_step"], "e2e %.4g" % j["e2e"]["value"], j.get("gpu_launches"), "frac %.3f" % r.get("frac", 0), r.get("kernel_ms"), "kept", r.get("rows_kept_fraction"))
    except Exception as e:
        print(n, "failed", e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
