#!/bin/bash
OUT=gpurun_out/exp5; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
for v in default minb8 minb6; do
  lib=spumoni_b200/lib/libspumoni_b200.so; [ $v != default ] && lib=spumoni_b200/lib/libspumoni_b200_$v.so
  SPUMONI_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  echo "== $v"; python - <<PY
import json
d=json.load(open("$OUT/bench_$v.json"))
print("u8 value %.2f G/s kernel %.2f ms e2e %.2f G/s | u32 value %.2f kernel %.2f ms e2e %.2f" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e9, d["variants"]["u32"]["value"]/1e9, d["variants"]["u32"]["kernel_ms"], d["variants"]["u32"]["e2e"]/1e9))
PY
  tail -2 $OUT/bench_$v.err
done
