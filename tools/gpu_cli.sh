#!/bin/bash
# CLI end to end on the c2 workload: `spumoni run -P -n -c`, FASTA in -> text out, the reference's timed phase.
# usage: bash tools/gpu_cli.sh <tag>
OUT=gpurun_out/${1:-cli}; mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > /dev/null 2>&1   # builds the c2 index
C=$(ls -d /tmp/spumoni_b200_cache/c2_*); REF=$C/idx
tools/bin/spumoni_mkindex reads --ref $REF.fa --n 1000000 --len 1000 --sub 0.06 --ins 0.02 --del 0.02 --neg-frac 0.1 --seed 4 --out $C/reads.fa > /dev/null
python -c "
import numpy as np; (np.arange(1,200,dtype=np.uint64)*7%11).tofile('/tmp/stats.bin')" && tools/bin/spumoni_mkindex nulldb --stats /tmp/stats.bin --out $REF.fa.pmlnulldb
for hf in 0 1; do for t in 1 4; do
  if [ $hf = 1 ]; then export SPUMONI_B200_HOST_FORMAT=1; else unset SPUMONI_B200_HOST_FORMAT; fi
  SPUMONI_B200_TIMING=1 spumoni_b200/bin/spumoni run -r $REF -p $C/reads.fa -P -n -c -t $t 2> $OUT/cli_hf${hf}_t$t.log
  echo "host_format=$hf -t $t: $(sed 's/\x1b\[[0-9;]*m//g' $OUT/cli_hf${hf}_t$t.log | grep -a 'busy\|^done' | tr '\n' ' ')  md5 $(md5sum < $C/reads.fa.pseudo_lengths | cut -c1-12)"
done; done
