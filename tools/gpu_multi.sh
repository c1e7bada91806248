#!/bin/bash
# Multi-GPU session (gpurun --gpus N).  usage: bash tools/gpu_multi.sh <tag> <N> [steps...]
#   steps: tests floor bench_c2 repl cli strong_c4s strong_c4s_1
set -u
export SPUMONI_B200_TIMING=1
TAG=${1:-multi}; N=${2:-2}; shift 2 || true
STEPS=${*:-tests floor bench_c2 repl cli}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > "$OUT/gpus.txt" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
nproc > "$OUT/nproc.txt"; free -g | head -2 >> "$OUT/nproc.txt"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
for s in $STEPS; do
  port=$((port + 1))
  case $s in
    tests)
      timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "replicated or two_gpus" > "$OUT/pytest_multi.log" 2>&1; tail -4 "$OUT/pytest_multi.log" ;;
    floor)
      for n in 1 $N; do
        port=$((port + 1))
        timeout 600 $TR --nproc-per-node $n --master-port $port tools/pcie_floor.py > "$OUT/pcie_floor_n$n.json" 2> "$OUT/pcie_floor_n$n.err"; cat "$OUT/pcie_floor_n$n.json"
      done ;;
    floor_all)
      for n in 1 2 4 8; do
        port=$((port + 1))
        timeout 600 $TR --nproc-per-node $n --master-port $port tools/pcie_floor.py > "$OUT/pcie_floor_n$n.json" 2> "$OUT/pcie_floor_n$n.err"; cat "$OUT/pcie_floor_n$n.json"
      done ;;
    bench_c2)
      timeout 1200 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --variants u8,packed_bits > "$OUT/bench_c2_n$N.json" 2> "$OUT/bench_c2_n$N.err"
      echo "bench c2 N=$N exit $?"; cat "$OUT/bench_c2_n$N.json"; tail -3 "$OUT/bench_c2_n$N.err" ;;
    repl)
      timeout 900 python tools/replicate_bench.py $N c2 ${REPL_REPS:-$N} > "$OUT/replicate_c2_n$N.json" 2> "$OUT/replicate_c2_n$N.err"; cat "$OUT/replicate_c2_n$N.json"; tail -2 "$OUT/replicate_c2_n$N.err" ;;
    cli)
      timeout 900 python bench.py --no-cpu-baseline --steps 2 --variants cli > "$OUT/cli_n$N.json" 2> "$OUT/cli_n$N.err"
      python -c "
import json; j=json.load(open('$OUT/cli_n$N.json')); print('cli', json.dumps(j.get('cli')))" ;;
    strong_c4s)
      SPB_BENCH_WORKLOAD=c4s timeout 1500 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 3 --warmup 3 --strong --variants u8,packed_bits > "$OUT/strong_c4s_n$N.json" 2> "$OUT/strong_c4s_n$N.err"
      echo "strong c4s N=$N exit $?"; cat "$OUT/strong_c4s_n$N.json"; tail -3 "$OUT/strong_c4s_n$N.err" ;;
    strong_c4s_1)
      SPB_BENCH_WORKLOAD=c4s timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 --strong --no-cpu-baseline --variants bits,packed_bits > "$OUT/strong_c4s_n1.json" 2> "$OUT/strong_c4s_n1.err"
      echo "strong c4s N=1 exit $?"; cat "$OUT/strong_c4s_n1.json"; tail -3 "$OUT/strong_c4s_n1.err" ;;
    *) echo "unknown step $s" ;;
  esac
done
ls -la "$OUT"
