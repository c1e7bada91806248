#!/bin/bash
# parity tests + PML bench for both row layouts + one full ncu capture; usage: bash tools/gpu_quick.sh <tag>
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
for rows in 8 16; do
SPUMONI_B200_ROWS=$rows timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); u=d['variants']['u32']; print('rows=$rows u8 value %.2f kernel %.2f ms e2e %.2f | u32 value %.2f e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9, u['value']/1e9, u['e2e']/1e9))"
done
timeout 900 python bench.py --mode ms --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MS value %.2f kernel %.2f ms e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"
bash tools/gpu_ncu.sh ${1:-quick} spb_move > /dev/null 2>&1
