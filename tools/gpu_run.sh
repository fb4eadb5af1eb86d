#!/bin/bash
# One parametrised GPU-box script (replaces the per-lease one-shot scripts): tools/gpu_run.sh TAG STEP [STEP ...]
#   steps: tests | bench[:workload] | ref | sanitize | launches[:workload] | ncu:KERNEL_REGEX[:workload] | py:SCRIPT
# Everything it writes goes to gpurun_out/<TAG>_*.
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for STEP in "$@"; do
  case $STEP in
    tests) python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log ;;
    bench*) WL=${STEP#bench}; WL=${WL#:}; WL=${WL:-c5}
      python bench.py --workload $WL > $OUT/${TAG}_bench_$WL.json 2> $OUT/${TAG}_bench_$WL.err; tail -c 600 $OUT/${TAG}_bench_$WL.json; tail -3 $OUT/${TAG}_bench_$WL.err ;;
    ref) python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cat $OUT/${TAG}_bench_ref.json ;;
    sanitize)
      for TOOL in memcheck racecheck; do
        timeout 900 compute-sanitizer --tool $TOOL --log-file $OUT/${TAG}_sanitizer_$TOOL.log \
          python tools/sanitize_case.py > $OUT/${TAG}_sanitizer_$TOOL.out 2>&1
        tail -4 $OUT/${TAG}_sanitizer_$TOOL.log
      done ;;
    launches*) WL=${STEP#launches}; WL=${WL#:}; WL=${WL:-c5}
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_$WL.csv \
        python bench.py --workload $WL --steps 2 --warmup 3 --headline-only --no-cpu-baseline > $OUT/${TAG}_launches_$WL.log 2>&1 ;;
    ncu:*) REST=${STEP#ncu:}; KREG=${REST%%:*}; WL=c2; [[ $REST == *:* ]] && WL=${REST#*:}
      ncu --set full --clock-control none --import-source on -k regex:$KREG -s 2 -c 1 -f -o $OUT/${TAG}_${KREG} \
        python bench.py --workload $WL --steps 1 --warmup 3 --headline-only --no-cpu-baseline > $OUT/${TAG}_ncu_${KREG}.log 2>&1
      tail -2 $OUT/${TAG}_ncu_${KREG}.log ;;
    py:*) S=${STEP#py:}; python $S > $OUT/${TAG}_$(basename $S .py).txt 2>&1; tail -40 $OUT/${TAG}_$(basename $S .py).txt ;;
  esac
done
