#!/bin/bash
# Round-2 experiment session 3 (c2): compact rows fetched as aligned pairs (SPB_ROW_PAIRS).
set -u
TAG=${1:-exp3}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
L=spumoni_b200/lib
P="$L/libspumoni_b200_v-SPB_ROW_PAIRS@1.so"; P9="$L/libspumoni_b200_v-SPB_ROW_PAIRS@1+SPB_MV_MINB_BITS@9.so"
SPUMONI_B200_LIB=$P timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_index_all_reads or bit_output or narrow or classifier or document or golden or edge" > "$OUT/pytest_pairs.log" 2>&1; tail -3 "$OUT/pytest_pairs.log"
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  python -c "
import json
j=json.load(open('$OUT/$name.json')); v=j.get('variants',{})
print('$name', 'value %.2f G kernel %.2f ms e2e %.2f G' % (j['value']/1e9, j['roofline']['kernel_ms'], j['e2e']['value']/1e9), ' '.join('%s: %.2f G' % (k, x.get('value',0)/1e9) for k,x in v.items() if isinstance(x,dict)))"
}
run c2_default -- --steps 5 --variants u8,pml_stats
run c2_pairs SPUMONI_B200_LIB=$P -- --steps 5 --variants u8,pml_stats
run c2_pairs9 SPUMONI_B200_LIB=$P9 -- --steps 5 --variants u8,pml_stats
