#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== pytest gpu diffusion"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k diffusion > gpurun_out/pytest_diff.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_diff.log
echo "== timing"; timeout 600 python tools/diffusion_time.py 16384 3 2>&1 | tail -3
timeout 600 python tools/diffusion_time.py 65536 2 2>&1 | tail -3
