#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
echo "== diag fwd+train"; timeout 600 python tools/diag1.py all 2>&1 | tail -40
echo "== racecheck (grid limited to 2 CTAs, 6 tiles)"; PSPDE_MAX_GRID=2 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/diag1.py small > gpurun_out/racecheck.log 2>&1; echo rc=$?; grep -v "^$" gpurun_out/racecheck.log | tail -25
echo "== memcheck"; PSPDE_MAX_GRID=2 timeout 900 compute-sanitizer --tool memcheck python tools/diag1.py small > gpurun_out/memcheck.log 2>&1; echo rc=$?; tail -8 gpurun_out/memcheck.log
