#!/bin/bash
# compute-sanitizer on the GPU box: memcheck (out-of-bounds / misaligned global and shared accesses) over a selection of the
# -m gpu parity tests that touches every kernel family once at small sizes (the tool serialises and instruments every launch).
# usage: tools/sanitize_gpu.sh [memcheck|racecheck|synccheck]   (writes gpurun_out/r02_sanitizer_<tool>.txt)
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out/r02_sanitizer_${TOOL}.txt
mkdir -p gpurun_out
SEL=${SANITIZE_SEL:-'test_tiled_vs_oracle or test_head_and_tail or (test_conv3x3_tc32_slab and shape0 and 1-1) or (test_conv2x2_tc32 and shape3) or test_blur or test_mri or test_multicoil or test_blurfft or test_tomography or test_dncnn_tc32_vs_oracle'}
FILES=${SANITIZE_FILES:-tests/test_gpu_radon_tiled.py tests/test_gpu_tc32.py tests/test_gpu_golden.py}
OUT=${SANITIZE_OUT:-$OUT}
timeout 900 compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 120 \
  python -m pytest $FILES -q -x -k "$SEL" > "$OUT" 2>&1
echo "exit=$?" >> "$OUT"
grep -c "Invalid\|Misaligned\|Race\|hazard" "$OUT" | sed 's/^/error lines: /' >> "$OUT"
tail -15 "$OUT"
