#!/bin/bash
# CLI text arena: page-locked (SPUMONI_B200_PINNED_TEXT=1) against pageable (default) on 10^9 bases of FASTA
# (4 Mb pangenome index: the index size only moves the kernel time, which is not what this measures).
# usage: bash tools/gpu_cli_pinned.sh <tag>
OUT=gpurun_out/${1:-clip}; mkdir -p $OUT
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -k "cli" -x -q 2>&1 | tail -3 | tee $OUT/pytest_cli.log
M=tools/bin/spumoni_mkindex; D=/tmp/clib; mkdir -p $D
$M genome --kind pangenome --len 4000000 --haps 8 --snp 0.001 --indel 0.0001 --seed 3 --out $D/raw.fa > /dev/null
$M build --input $D/raw.fa --prefix $D/idx --pml > /dev/null
$M reads --ref $D/idx.fa --n 1000000 --len 1000 --sub 0.06 --ins 0.02 --del 0.02 --neg-frac 0.1 --seed 4 --out $D/reads.fa > /dev/null
python -c "
import numpy as np; (np.arange(1,200,dtype=np.uint64)*7%11).tofile('/tmp/stats.bin')" && $M nulldb --stats /tmp/stats.bin --out $D/idx.fa.pmlnulldb
for mode in pinned pageable pinned pageable; do
  if [ $mode = pinned ]; then export SPUMONI_B200_PINNED_TEXT=1; else unset SPUMONI_B200_PINNED_TEXT; fi
  SPUMONI_B200_TIMING=1 timeout 60 spumoni_b200/bin/spumoni run -r $D/idx -p $D/reads.fa -P -n -c -t 1 2> $OUT/cli_$mode.log
  echo "$mode: $(sed 's/\x1b\[[0-9;]*m//g' $OUT/cli_$mode.log | grep -a 'busy\|processing' | tr '\n' ' ')  md5 $(md5sum < $D/reads.fa.pseudo_lengths | cut -c1-12)" | tee -a $OUT/summary.txt
done
