#!/bin/bash
OUT=gpurun_out/exp13; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('u8 value %.2f kernel %.2f ms e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"
timeout 900 python bench.py --mode ms --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('MS value %.2f kernel %.2f ms e2e %.2f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['e2e']['value']/1e9))"
# CLI end to end on the c2 workload: 1 M x 1 kb reads as FASTA, wall clock of the "processing" phase
C=$(ls -d /tmp/spumoni_b200_cache/c2_*); REF=$C/idx
tools/bin/spumoni_mkindex reads --ref $REF.fa --n 1000000 --len 1000 --sub 0.06 --ins 0.02 --del 0.02 --neg-frac 0.1 --seed 4 --out $C/reads.fa > /dev/null
tools/bin/spumoni_mkindex nulldb --stats <(python -c "import numpy as np,sys; sys.stdout.buffer.write((np.arange(1,200,dtype=np.uint64)*7%11).tobytes())") --out $REF.fa.pmlnulldb 2>/dev/null || python -c "
import numpy as np; (np.arange(1,200,dtype=np.uint64)*7%11).tofile('/tmp/stats.bin')" && tools/bin/spumoni_mkindex nulldb --stats /tmp/stats.bin --out $REF.fa.pmlnulldb
ls -la $C | head -20
for t in 1 16; do S=$(date +%s.%N); spumoni_b200/bin/spumoni run -r $REF -p $C/reads.fa -P -n -c -t $t 2> $OUT/cli_t$t.log; grep -a "done\|finished" $OUT/cli_t$t.log; echo "wall $(echo "$(date +%s.%N) - $S" | bc) s (-t $t)"; done
head -c 300 $C/reads.fa.pseudo_lengths; echo; head -3 $C/reads.fa.report; ls -la $C/reads.fa*
head -n 20000 $C/reads.fa > $C/sample.fa
oracle/_build/spumoni_oracle run -r $REF -p $C/sample.fa -P -n -c -t $(nproc) 2> $OUT/oracle_cli.log; grep -a "oracle" $OUT/oracle_cli.log
head -n 20000 $C/reads.fa.pseudo_lengths | cmp - $C/sample.fa.pseudo_lengths && echo "CLI lengths identical to the oracle CLI on the first 10000 reads (order: -t 1 only)"
