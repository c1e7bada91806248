#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (PML headline + MS secondary), reference arm, ncu launch
# list, one `ncu --set full` capture of the dominant kernel.  Usage (from the repo root, on the GPU box):
#   bash tools/gpu_round.sh <tag>
# Everything that should come back is written under gpurun_out/<tag>/.
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/nproc.txt"

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
  timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1
  tail -2 "$OUT/smoke.log"
fi

# bench (device-timed value, e2e, cpu_baseline); extra args from $BENCH_ARGS
timeout 1500 python bench.py ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"

if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
  cat "$OUT/bench_reference.json"
fi

if [ "${SKIP_MS:-0}" != "1" ]; then
  timeout 1500 python bench.py --mode ms ${BENCH_ARGS:-} > "$OUT/bench_ms.json" 2> "$OUT/bench_ms.err"
  echo "bench ms exit $?"; cat "$OUT/bench_ms.json"; tail -3 "$OUT/bench_ms.err"
fi

if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of the same command (cold-cache, serialised: shares, not absolutes)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/launches_bench.log" 2>&1
  # one full capture of the dominant kernel
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-spb_move} -s 3 -c 1 \
      -f -o "$OUT/prof" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_bench.log" 2>&1
  if [ "${SKIP_MS:-0}" != "1" ]; then
    timeout 1500 ncu --set full --clock-control none --import-source on -k regex:spb_m -s 6 -c 2 \
        -f -o "$OUT/prof_ms" python bench.py --mode ms --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_ms_bench.log" 2>&1
  fi
  ls -la "$OUT"
fi
