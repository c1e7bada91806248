#!/bin/bash
OUT=gpurun_out/exp16; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py --mode ms --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_ms.json | python -c "import json,sys; d=json.load(sys.stdin); u=d['variants']['u32']; print('MS u8 value %.2f kernel %.2f ms e2e %.2f | u32 value %.2f e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9, u['value']/1e9, u['e2e']/1e9))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:spb_m -c 12 --csv python bench.py --mode ms --steps 1 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | grep -a "spb_" | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('PML u8 value %.2f kernel %.2f ms e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"
