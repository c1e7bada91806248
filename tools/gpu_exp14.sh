#!/bin/bash
OUT=gpurun_out/exp14; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -6 $OUT/pytest_gpu.log
timeout 900 python bench.py --mode ms --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_ms.json | python -c "import json,sys; d=json.load(sys.stdin); u=d['variants']['u32']; print('MS u8 value %.2f kernel %.2f ms e2e %.2f | u32 value %.2f e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9, u['value']/1e9, u['e2e']/1e9))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spb_m -c 12 --csv python bench.py --mode ms --steps 1 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | grep -a "spb_" | awk -F'","' '{print $5, $NF}' | tail -4
C=$(ls -d /tmp/spumoni_b200_cache/c2_*); REF=$C/idx
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > /dev/null 2>&1   # builds the PML index too
tools/bin/spumoni_mkindex reads --ref $REF.fa --n 1000000 --len 1000 --sub 0.06 --ins 0.02 --del 0.02 --neg-frac 0.1 --seed 4 --out $C/reads.fa > /dev/null
python -c "
import numpy as np; (np.arange(1,200,dtype=np.uint64)*7%11).tofile('/tmp/stats.bin')" && tools/bin/spumoni_mkindex nulldb --stats /tmp/stats.bin --out $REF.fa.pmlnulldb
for t in 1 4 16; do spumoni_b200/bin/spumoni run -r $REF -p $C/reads.fa -P -n -c -t $t 2> $OUT/cli_t$t.log; echo "-t $t: $(grep -a 'processing' $OUT/cli_t$t.log | sed 's/\x1b\[[0-9;]*m//g')"; done
md5sum $C/reads.fa.pseudo_lengths $C/reads.fa.report
head -n 4000 $C/reads.fa > $C/sample.fa
oracle/_build/spumoni_oracle run -r $REF -p $C/sample.fa -P -n -c -t 1 2> $OUT/oracle_cli.log; grep -a "oracle" $OUT/oracle_cli.log
head -n 4000 $C/reads.fa.pseudo_lengths | cmp - $C/sample.fa.pseudo_lengths && echo "CLI lengths identical to the oracle CLI (-t 1) on the first 2000 reads"
head -n 2001 $C/reads.fa.report | cmp - $C/sample.fa.report && echo "CLI report identical to the oracle CLI (-t 1) on the first 2000 reads"
