#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (bit output, packed input, 32-byte rows, text path for MS / -d, GPU builder).
set -u
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
K='bit_output or bits_device or narrow or mixed_lengths or on_device_text or golden_vectors or edge_cases or document_numbers or synthetic_cuda'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_zz_synthetic_index.py -m gpu -x -q -k "$K" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?" >> "$OUT/memcheck.log"; tail -6 "$OUT/memcheck.log"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_build.py -m gpu -x -q -k "repeats or binary or unary or rejects" > "$OUT/memcheck_build.log" 2>&1; echo "memcheck exit $?" >> "$OUT/memcheck_build.log"; tail -5 "$OUT/memcheck_build.log"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_output_and_packed or bit_output_many or on_device_text or cli_fast_path_equals" > "$OUT/racecheck.log" 2>&1; echo "racecheck exit $?" >> "$OUT/racecheck.log"; tail -5 "$OUT/racecheck.log"
