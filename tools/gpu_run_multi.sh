#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== phase profile"; timeout 300 python tools/phase_profile.py c2 2>&1 | tail -12 | tee gpurun_out/phase_profile_c2.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 600 gpurun_out/bench_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -c 1500 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json","gpurun_out/bench_n2.json"):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, "value %.4g ms/step %.2f e2e %.4g launches %s kernels %s frac %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["kernel_ms"] if d.get("roofline") else None, d["roofline"]["frac"] if d.get("roofline") else None))
    except Exception as e: print(f, "ERR", e)
PY
echo "== other workloads (N=1)"; for w in c1 c3 c3re; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/bench_$w.json').read().splitlines() if l.startswith('{')][-1]); print('$w', 'value %.4g ms/step %.3f loss %s' % (d['value'], d['ms_per_step'], d['final_loss']))
except Exception as e: print('$w ERR', e, open('gpurun_out/bench_$w.err').read()[-800:])
"; done
