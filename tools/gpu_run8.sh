#!/bin/bash
# f4 (elliptic) bring-up on one B200: the new GPU tests first, then the whole GPU suite and a C2 bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== pytest elliptic"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "elliptic" > gpurun_out/pytest_ell.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_ell.log
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?; tail -c 300 gpurun_out/bench_n1.err; head -c 600 gpurun_out/bench_n1.json
