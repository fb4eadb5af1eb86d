#!/bin/bash
# evidence for the single-rollout step: smoke, all GPU tests, launch list and full ncu capture of its two kernels
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== ncu launches c2"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1d_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1; echo rc=$?
echo "== ncu full c2 (forward that keeps its rows + gradient kernel, in the bench pipeline)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"^grad_tc_kernel|^rollout_tc_fwd_kernel" -s 6 -c 2 -o gpurun_out/prof_r1d_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_f2.log
ls -la gpurun_out/*.ncu-rep
