#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench c4"; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo rc=$?; tail -c 300 gpurun_out/bench_c4.err; cut -c1-600 gpurun_out/bench_c4.json
echo "== ncu launches c4"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv python tools/diffusion_time.py 16384 1 > gpurun_out/ncu_l.log 2>&1; echo rc=$?
echo "== ncu full c4"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:diffusion_kernel -c 2 -o gpurun_out/prof_c4 python tools/diffusion_time.py 16384 0 > gpurun_out/ncu_f.log 2>&1; echo rc=$?; tail -3 gpurun_out/ncu_f.log
