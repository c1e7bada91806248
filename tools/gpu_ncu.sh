#!/bin/bash
# one `ncu --set full` capture of the dominant kernel of bench.py; usage: bash tools/gpu_ncu.sh <tag> [kernel regex] [extra bench args]
TAG=${1:-ncu}; K=${2:-spb_move}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-variants "$@" > $OUT/prof_bench.log 2>&1
tail -3 $OUT/prof_bench.log; ls -la $OUT
