#!/bin/bash
# Round-2 experiment session: kernel occupancy variants on c2, balancing / cache-policy variants on c4s, the CLI fast path.
# usage: bash tools/gpu_exp1.sh <tag>
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-exp1}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
L=spumoni_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_lengths or bit_output or replicated or narrow" > "$OUT/pytest_sel.log" 2>&1; tail -3 "$OUT/pytest_sel.log"
run() {  # name, env..., then bench args after --
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
          " ".join("%s: %.2f G / e2e %.2f G" % (k, x.get("value", 0) / 1e9, x.get("e2e", 0) / 1e9) for k, x in v.items() if isinstance(x, dict) and "e2e" in x),
          ("cli %.2f G (%s)" % (j["cli"]["bases_per_s"] / 1e9, j["cli"].get("stages", "")) if j.get("cli") else ""))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
# ---- c2: occupancy variants (main = u8 kernel, variant = bits kernel)
run c2_default -- --steps 5 --variants bits
run c2_bits10 SPUMONI_B200_LIB=$L/libspumoni_b200_v-SPB_MV_MINB_BITS@10.so -- --steps 5 --variants bits
run c2_bits12 SPUMONI_B200_LIB=$L/libspumoni_b200_v-SPB_MV_MINB_BITS@12.so -- --steps 5 --variants bits
run c2_stage32_10 "SPUMONI_B200_LIB=$L/libspumoni_b200_v-SPB_STAGE@32+SPB_MV_MINB@10.so" -- --steps 5 --variants bits
run c2_l2fetch32 SPUMONI_B200_L2_FETCH=32 -- --steps 5 --variants bits
# ---- c2: the CLI (fast path) and its sequential path
run c2_cli -- --steps 3 --variants cli
run c2_cli_seq SPUMONI_B200_SEQUENTIAL=1 -- --steps 3 --variants cli
# ---- c4s: balancing and cache policy
run c4s_default SPB_BENCH_WORKLOAD=c4s -- --steps 3 --variants bits
run c4s_bal42 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_BALANCE=4,2,4093 -- --steps 3 --no-variants
run c4s_bal21 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_BALANCE=2,1,4093 -- --steps 3 --no-variants
run c4s_bal31 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_BALANCE=3,1,4093 -- --steps 3 --no-variants
run c4s_rowld0 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_ROW_LD=0 -- --steps 3 --no-variants
run c4s_l2fetch32 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_L2_FETCH=32 -- --steps 3 --no-variants
run c4s_rows16 SPB_BENCH_WORKLOAD=c4s SPUMONI_B200_ROWS=16 -- --steps 3 --no-variants
run c4s_stage32_10 SPB_BENCH_WORKLOAD=c4s "SPUMONI_B200_LIB=$L/libspumoni_b200_v-SPB_STAGE@32+SPB_MV_MINB@10.so" -- --steps 3 --variants bits
grep -h "rows)" "$OUT"/c4s_*.err | sort | uniq -c | head
