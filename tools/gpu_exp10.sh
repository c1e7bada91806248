#!/bin/bash
OUT=gpurun_out/exp10; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
for rows in 8 16; do
  SPUMONI_B200_ROWS=$rows timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_rows$rows.json 2> $OUT/bench_rows$rows.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_rows$rows.json")); u=d["variants"]["u32"]
print("rows=$rows u8 value %6.2f G/s kernel %6.2f ms e2e %6.2f G/s | u32 value %6.2f kernel %6.2f ms e2e %6.2f" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e9, u["value"]/1e9, u["kernel_ms"], u["e2e"]/1e9))
PY
done
SPUMONI_B200_ROWS=8 timeout 900 python bench.py --mode ms --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MS rows=8 value %.2f e2e %.2f' % (d['value']/1e9, d['e2e']['value']/1e9))"
bash tools/gpu_ncu.sh exp10 spb_move
