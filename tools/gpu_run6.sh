#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== pytest tc"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "tc_" > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_tc.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 400 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n1.json").read().splitlines() if l.startswith("{")][-1])
print("value %.4g ms/step %.2f e2e %.4g kernels %s frac %s step %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["step"]["frac"]))
PY
