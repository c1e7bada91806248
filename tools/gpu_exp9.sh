#!/bin/bash
OUT=gpurun_out/exp9; mkdir -p $OUT
bash tools/gpu_variants.sh exp9 --no-variants
echo "--- row_ld 1 / 2 on default"
for ld in 1 2; do SPUMONI_B200_ROW_LD=$ld python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ld=$ld value %.2f e2e %.2f' % (d['value']/1e9, d['e2e']['value']/1e9))"; done
echo "--- e2e pipeline knobs (default lib)"
for cfg in "3 32" "8 32" "12 16" "16 16" "8 64" "16 8" "6 128"; do set -- $cfg
  SPUMONI_B200_SLOTS=$1 SPUMONI_B200_CHUNK_MB=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('slots=$1 chunk=$2 value %.2f e2e %.2f' % (d['value']/1e9, d['e2e']['value']/1e9))"
done
SPUMONI_B200_SLOTS=8 SPUMONI_B200_CHUNK_MB=32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --length-bytes 4 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('u32 slots=8 chunk=32 value %.2f e2e %.2f' % (d['value']/1e9, d['e2e']['value']/1e9))"
SPUMONI_B200_SLOTS=8 SPUMONI_B200_CHUNK_MB=32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --length-bytes 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('u16 slots=8 chunk=32 value %.2f e2e %.2f' % (d['value']/1e9, d['e2e']['value']/1e9))"
