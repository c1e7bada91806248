#!/bin/bash
# c4r (real-text 1e9-character index built on the GPU) bench + ncu, read-length sweeps, new CLI tests.  usage: bash tools/gpu_r2g.sh <tag>
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-r2g}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli_fast or hands_over or cli_outputs or on_device_text" > "$OUT/pytest_cli.log" 2>&1; tail -4 "$OUT/pytest_cli.log"
SPB_BENCH_WORKLOAD=c4r timeout 1500 python bench.py --steps 5 --variants u8,bits > "$OUT/bench_c4r.json" 2> "$OUT/bench_c4r.err"
echo "bench c4r exit $?"; cat "$OUT/bench_c4r.json"; tail -8 "$OUT/bench_c4r.err"
SPB_BENCH_WORKLOAD=c4r timeout 900 ncu --set full --clock-control none --import-source on -k regex:spb_move -s 3 -c 1 \
    -f -o "$OUT/prof_c4r" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_c4r.log" 2>&1
timeout 900 python tools/sweep_bench.py > "$OUT/sweep_c2.json" 2> "$OUT/sweep_c2.err"; cat "$OUT/sweep_c2.json"; tail -2 "$OUT/sweep_c2.err"
SPB_BENCH_WORKLOAD=c4s timeout 900 python tools/sweep_bench.py > "$OUT/sweep_c4s.json" 2> "$OUT/sweep_c4s.err"; cat "$OUT/sweep_c4s.json"; tail -2 "$OUT/sweep_c4s.err"
ls -la "$OUT"
