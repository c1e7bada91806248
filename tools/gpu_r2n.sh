#!/bin/bash
# smoke + final numbers of the two config-4 stand-ins with the final kernels.  usage: bash tools/gpu_r2n.sh <tag>
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-r2n}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; tail -1 "$OUT/smoke.log"
SPB_BENCH_WORKLOAD=c4r timeout 1200 python bench.py --steps 5 --variants u8,bits,pml_stats > "$OUT/bench_c4r.json" 2> "$OUT/bench_c4r.err"; echo "c4r exit $?"; cat "$OUT/bench_c4r.json" | cut -c1-400
SPB_BENCH_WORKLOAD=c4s timeout 1200 python bench.py --steps 5 --no-cpu-baseline --variants u8,bits,pml_stats > "$OUT/bench_c4s.json" 2> "$OUT/bench_c4s.err"; echo "c4s exit $?"; cat "$OUT/bench_c4s.json" | cut -c1-400
