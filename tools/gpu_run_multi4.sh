#!/bin/bash
# 2-GPU check of the single-rollout step: NCCL tests + weak-scaling bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; echo rc=$?; tail -c 300 gpurun_out/scale_n2.err
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo rc=$?
python - <<'PY'
import json
for f in ("scale_n1","scale_n2"):
    d=json.loads([l for l in open("gpurun_out/%s.json"%f).read().splitlines() if l.startswith("{")][-1])
    print(f, "value %.4g ms/step %.2f n_gpus %s e2e %.4g" % (d["value"], d["ms_per_step"], d.get("n_gpus"), d["e2e"]["value"]))
PY
