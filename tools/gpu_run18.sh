#!/bin/bash
# single-rollout step: parity tests, then the C2 bench (default = single rollout) and the two-rollout step for comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_rollout or checkpointed or mn_major or building_block" 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_single.json 2> gpurun_out/bench_c2_single.err; tail -c 3000 gpurun_out/bench_c2_single.json; tail -3 gpurun_out/bench_c2_single.err
PSPDE_FWD_CKPT_MAX_GB=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2_two.json 2>/dev/null; python - <<'PY'
import json
for n in ("single", "two"):
    try:
        j = json.loads(open("gpurun_out/bench_c2_%s.json" % n).read().strip().splitlines()[-1])
        print(n, j["value"], j["ms_per_step"], j.get("e2e"), j.get("gpu_launches"), j.get("roofline", {}).get("frac"))
    except Exception as e:
        print(n, "failed", e)
PY
