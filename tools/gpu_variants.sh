#!/bin/bash
# bench every development variant library in spumoni_b200/lib (plus the default); usage: bash tools/gpu_variants.sh <tag> [bench args]
TAG=${1:-variants}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in spumoni_b200/lib/libspumoni_b200.so spumoni_b200/lib/libspumoni_b200_v-*.so; do
  v=$(basename $lib .so); v=${v#libspumoni_b200}; v=${v:-_default}
  SPUMONI_B200_LIB=$PWD/$lib timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench$v.json 2> $OUT/bench$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench$v.json"))
    u=d["variants"].get("u32")
    print("%-60s u8 value %6.2f G/s kernel %6.2f ms e2e %6.2f G/s" % ("$v", d["value"]/1e9, d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e9) + (" | u32 value %6.2f kernel %6.2f ms e2e %6.2f" % (u["value"]/1e9, u["kernel_ms"], u["e2e"]/1e9) if u else ""))
except Exception as e:
    print("$v", "FAILED", e); print(open("$OUT/bench$v.err").read()[-500:])
PY
done
