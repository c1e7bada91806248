#!/bin/bash
OUT=gpurun_out/exp7; mkdir -p $OUT
for v in default minb10; do for ld in 0 2; do
  lib=spumoni_b200/lib/libspumoni_b200.so; [ $v != default ] && lib=spumoni_b200/lib/libspumoni_b200_$v.so
  SPUMONI_B200_ROW_LD=$ld SPUMONI_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${v}_ld$ld.json 2> $OUT/bench_${v}_ld$ld.err
  echo "== $v ld=$ld"; python - <<PY
import json
d=json.load(open("$OUT/bench_${v}_ld$ld.json"))
print("u8 value %.2f G/s kernel %.2f ms e2e %.2f G/s | u32 value %.2f kernel %.2f ms e2e %.2f" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e9, d["variants"]["u32"]["value"]/1e9, d["variants"]["u32"]["kernel_ms"], d["variants"]["u32"]["e2e"]/1e9))
PY
  tail -1 $OUT/bench_${v}_ld$ld.err
done; done
