#!/bin/bash
# experiment: L2 fetch granularity / table size / load flavour / occupancy (gather_bench) and the real kernel
OUT=gpurun_out/exp1; mkdir -p $OUT
G=tools/bin/gather_bench
{
for mib in 32 64 96 128 160 256 640 2048; do for gran in 0 32; do $G $mib $gran 0 1280 1500; done; done
for gran in 64 128; do $G 640 $gran 0 1280 1500; done
for fl in 1 2 3 4; do for gran in 0 32; do $G 640 $gran $fl 1280 1500; done; done
for fl in 3; do $G 96 32 $fl 1280 1500; $G 160 32 $fl 1280 1500; done
for t in 256 512 768 1024 1536 2048; do $G 640 32 0 $t 1500; done
} > $OUT/gather.txt 2>&1
cat $OUT/gather.txt
for g in 0 32 64; do
  SPUMONI_B200_L2_FETCH=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_l2fetch_$g.json 2> $OUT/bench_l2fetch_$g.err
  python -c "import json;d=json.load(open('$OUT/bench_l2fetch_$g.json'));print('l2fetch',$g,'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
