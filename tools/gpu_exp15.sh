#!/bin/bash
OUT=gpurun_out/exp15; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -k "cli" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > /dev/null 2>&1
C=$(ls -d /tmp/spumoni_b200_cache/c2_*); REF=$C/idx
tools/bin/spumoni_mkindex reads --ref $REF.fa --n 1000000 --len 1000 --sub 0.06 --ins 0.02 --del 0.02 --neg-frac 0.1 --seed 4 --out $C/reads.fa > /dev/null
python -c "
import numpy as np; (np.arange(1,200,dtype=np.uint64)*7%11).tofile('/tmp/stats.bin')" && tools/bin/spumoni_mkindex nulldb --stats /tmp/stats.bin --out $REF.fa.pmlnulldb
for t in 1 4 16; do SPUMONI_B200_TIMING=1 spumoni_b200/bin/spumoni run -r $REF -p $C/reads.fa -P -n -c -t $t 2> $OUT/cli_t$t.log; echo "-t $t: $(grep -a 'processing\|timing' $OUT/cli_t$t.log | sed 's/\x1b\[[0-9;]*m//g')"; done
md5sum $C/reads.fa.pseudo_lengths $C/reads.fa.report
SPUMONI_B200_TIMING=1 spumoni_b200/bin/spumoni run -r $REF -p $C/reads.fa -P -n -t 16 2>&1 | grep -a 'processing\|timing' | sed 's/\x1b\[[0-9;]*m//g'
