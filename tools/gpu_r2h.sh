#!/bin/bash
# Final verification session of the round: full GPU suite, smoke, both bench arms (default flags), ncu launch list + full capture, sweep.
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-r2h}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1; nproc > "$OUT/nproc.txt"
timeout 1800 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; tail -6 "$OUT/pytest_gpu.log"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; cat "$OUT/bench_reference.json"
timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"; cat "$OUT/bench.json"; tail -6 "$OUT/bench.err"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/launches_bench.log" 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:spb_move -s 3 -c 1 -f -o "$OUT/prof" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_bench.log" 2>&1
timeout 900 python tools/sweep_bench.py > "$OUT/sweep_c2.json" 2> "$OUT/sweep_c2.err"; cat "$OUT/sweep_c2.json"; tail -2 "$OUT/sweep_c2.err"
ls -la "$OUT"
