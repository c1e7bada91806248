#!/bin/bash
# FASTQ on the CLI's fast path: the CLI tests, then `spumoni run` on configs[1]'s reads as FASTA and as FASTQ.  usage: bash tools/gpu_r2r.sh <tag>
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-r2r}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fastq" > "$OUT/pytest_cli.log" 2>&1; echo "pytest exit $?"; tail -3 "$OUT/pytest_cli.log"
timeout 150 python bench.py --steps 3 --no-cpu-baseline --variants cli,cli_fastq > "$OUT/bench_cli.json" 2> "$OUT/bench_cli.err"; echo "bench exit $?"
python - <<'PY'
import json,sys
try:
    j=json.loads(open("gpurun_out/%s/bench_cli.json" % (sys.argv[1] if len(sys.argv)>1 else "r2r")).read().strip().splitlines()[-1])
    print(json.dumps(j.get("cli"))[:1500])
except Exception as e: print("no json", e)
PY
tail -5 "$OUT/bench_cli.err" | cut -c1-300
