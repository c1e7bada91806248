#!/bin/bash
# builder tests, full GPU suite, CLI write modes on c2, file-system write floor.  usage: bash tools/gpu_r2d.sh <tag>
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-r2d}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_build.py -m gpu -x -q > "$OUT/pytest_build.log" 2>&1; tail -25 "$OUT/pytest_build.log"
timeout 1800 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1; tail -6 "$OUT/pytest_gpu.log"
for wm in pwrite mmap write; do
  SPUMONI_B200_WRITE=$wm timeout 600 python bench.py --no-cpu-baseline --steps 3 --variants cli > "$OUT/cli_$wm.json" 2> "$OUT/cli_$wm.err"
  python -c "
import json,sys
j=json.load(open('$OUT/cli_$wm.json')); c=j.get('cli') or {}
print('$wm', 'cli %.2f Gbases/s' % (c.get('bases_per_s',0)/1e9), c.get('phase_s'), c.get('stages'))"
done
C=$(ls -d /tmp/spumoni_b200_cache/c2_*)
tools/bin/fs_write_bench $C/wb.out 2700 16 > "$OUT/fs_write_bench.txt" 2>&1; cat "$OUT/fs_write_bench.txt"
df -T $C | tail -1 >> "$OUT/fs_write_bench.txt"
# the c2 text through the device builder against mkindex's intermediates
python - <<'PY' > "$OUT/build_c2.txt" 2>&1
import os, sys, time, glob, subprocess, numpy as np
sys.path.insert(0, '.')
import spumoni_b200 as sp
C = glob.glob('/tmp/spumoni_b200_cache/c2_*')[0]
T = np.frombuffer("".join(l.strip() for l in open(C + '/idx.fa') if not l.startswith('>')).encode() + b'\0', dtype=np.uint8)
t0 = time.time(); P = sp.build_parts_device(T, 0); t1 = time.time()
print('device build of c2 text: n=%d r=%d in %.2f s (incl. H2D/D2H)' % (P['n'], P['r'], t1 - t0))
t0 = time.time(); P = sp.build_parts_device(T, 0); print('second call %.2f s' % (time.time() - t0))
idx = sp.Index.open(C + '/idx.fa', sp.MODE_PML)
print('index on disk: n, r =', idx.get_bwt_stats())
PY
cat "$OUT/build_c2.txt"
