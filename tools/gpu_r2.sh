#!/bin/bash
# Round-2 GPU session (one gpurun call).  usage: bash tools/gpu_r2.sh <tag> [steps...]
#   steps: tests smoke bench bench_c4s ncu ncu_c4s ref ms cli   (default: tests smoke bench)
set -u
TAG=${1:-r2}; shift || true
STEPS=${*:-tests smoke bench}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/nproc.txt"
for s in $STEPS; do
  case $s in
    tests)
      timeout 1800 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -15 "$OUT/pytest_gpu.log" ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log" ;;
    bench)
      timeout 1500 python bench.py ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"
      echo "bench exit $?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err" ;;
    bench_c4s)
      SPB_BENCH_WORKLOAD=c4s timeout 1500 python bench.py ${BENCH_ARGS:-} > "$OUT/bench_c4s.json" 2> "$OUT/bench_c4s.err"
      echo "bench c4s exit $?"; cat "$OUT/bench_c4s.json"; tail -5 "$OUT/bench_c4s.err" ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
      cat "$OUT/bench_reference.json" ;;
    ncu|ncu_c4s)
      W=c2; [ $s = ncu_c4s ] && W=c4s
      SPB_BENCH_WORKLOAD=$W timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
          --log-file "$OUT/launches_$W.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/launches_$W.log" 2>&1
      SPB_BENCH_WORKLOAD=$W timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-spb_move} -s 3 -c 1 \
          -f -o "$OUT/prof_$W" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_$W.log" 2>&1 ;;
    *) echo "unknown step $s" ;;
  esac
done
ls -la "$OUT"
