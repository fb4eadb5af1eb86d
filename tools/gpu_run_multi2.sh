#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L | head -8
NG=$(nvidia-smi -L | wc -l)
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 2>&1 | tail -4
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "== bench N=$n"
    if [ $n -eq 1 ]; then timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
    echo rc=$?; tail -c 400 gpurun_out/scale_n$n.err
  fi
done
python - <<'PY'
import json, glob
base=None
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open("gpurun_out/scale_n%d.json"%n).read().splitlines() if l.startswith("{")][-1])
        if base is None: base=d["value"]
        print("N=%d value %.4g ms/step %.2f e2e %.4g scaling %.2fx" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["value"]/base))
    except Exception as e: print(n, "ERR", e)
PY
