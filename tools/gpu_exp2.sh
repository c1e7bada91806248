#!/bin/bash
# Round-2 experiment session 2 (c2): asynchronous read staging (cp.async double buffer), L2 persisting window over the move rows.
set -u
TAG=${1:-exp2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
L=spumoni_b200/lib
A="$L/libspumoni_b200_v-SPB_ASYNC_READS@1.so"; A9="$L/libspumoni_b200_v-SPB_ASYNC_READS@1+SPB_MV_MINB_BITS@9.so"
SPUMONI_B200_LIB=$A timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_index_all_reads or long_and_mixed or bit_output or narrow or mixed_lengths or classifier or document" > "$OUT/pytest_async.log" 2>&1; tail -3 "$OUT/pytest_async.log"
SPUMONI_B200_L2_PERSIST=100 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_index_all_reads or bit_output or narrow or ms_pointers32" > "$OUT/pytest_persist.log" 2>&1; tail -3 "$OUT/pytest_persist.log"
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python bench.py --no-cpu-baseline "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    v = j.get("variants", {})
    print(sys.argv[2], "value %.2f G  kernel %.2f ms  e2e %.2f G" % (j["value"] / 1e9, j["roofline"]["kernel_ms"], j["e2e"]["value"] / 1e9),
          " ".join("%s: %.2f G / e2e %.2f G" % (k, x.get("value", 0) / 1e9, x.get("e2e", 0) / 1e9) for k, x in v.items() if isinstance(x, dict) and "e2e" in x))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run c2_default -- --steps 5 --variants u8
run c2_async SPUMONI_B200_LIB=$A -- --steps 5 --variants u8
run c2_async9 SPUMONI_B200_LIB=$A9 -- --steps 5 --variants u8
run c2_persist100 SPUMONI_B200_L2_PERSIST=100 -- --steps 5 --variants u8
run c2_persist60 SPUMONI_B200_L2_PERSIST=60 -- --steps 5 --variants u8
run c2_persist100_async9 SPUMONI_B200_L2_PERSIST=100 SPUMONI_B200_LIB=$A9 -- --steps 5 --variants u8
run c2_ms_default -- --mode ms --steps 4 --no-variants
run c2_ms_persist100 SPUMONI_B200_L2_PERSIST=100 -- --mode ms --steps 4 --no-variants
run c2_ms_async SPUMONI_B200_LIB=$A -- --mode ms --steps 4 --no-variants
