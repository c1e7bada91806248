#!/bin/bash
# MS (configs[2]) A/B: where the per-row SA samples live.  usage: bash tools/gpu_ms.sh <tag>
set -u
TAG=${1:-ms}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ms_sample_layouts or ms_pointers32 or small_index_all_reads" > "$OUT/pytest_ms.log" 2>&1; tail -3 "$OUT/pytest_ms.log"
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline --mode ms --no-variants "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  python -c "
import json,sys
j=json.load(open('$OUT/$name.json')); print('$name', 'value %.2f G kernel %.2f ms e2e %.2f G' % (j['value']/1e9, j['roofline']['kernel_ms'], j['e2e']['value']/1e9))"
}
run ms_inline -- --steps 4
run ms_split32 SPUMONI_B200_MSAMP=split32 -- --steps 4
run ms_split32_persist SPUMONI_B200_MSAMP=split32 SPUMONI_B200_MS_PERSIST=1 -- --steps 4
run ms_wide SPUMONI_B200_MSAMP=wide -- --steps 4
SPUMONI_B200_MSAMP=split32 SPUMONI_B200_MS_PERSIST=1 timeout 900 ncu --set full --clock-control none -k regex:spb_m -s 6 -c 2 -f -o "$OUT/prof_ms_split32" python bench.py --mode ms --steps 1 --warmup 3 --no-cpu-baseline --no-variants > "$OUT/prof_ms.log" 2>&1
