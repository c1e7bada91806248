#!/bin/bash
# experiment: move-row kernels -- parity tests, then bench with row-load flavours / output widths
OUT=gpurun_out/exp3; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
for ld in 0 1; do
  SPUMONI_B200_ROW_LD=$ld timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_ld$ld.json 2> $OUT/bench_ld$ld.err
  echo "== row_ld $ld"; cat $OUT/bench_ld$ld.json; tail -3 $OUT/bench_ld$ld.err
done
