#!/bin/bash
# One gpurun call that produces everything a round needs from the B200: the GPU test suite, smoke(), the bench lines of the four
# configurations (+ the reference arm), the operator / layer micro-benchmarks, an ncu launch list of bench steps and
# `--set full` captures of the named kernels.
#   usage (from the repo root, in the authoring container):
#     /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02_final "conv_tc32_slab"'
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/ (tools/ncu_extract.py condenses ncu files).
set -u
TAG=${1:-sess}
KERNELS=${2:-}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; tail -2 $OUT/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; cut -c1-200 $OUT/${TAG}_bench_reference.json
for c in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --lean > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; cut -c1-160 $OUT/${TAG}_bench_$c.json
done
timeout 240 python tools/op_bench.py mri tomo blur mcmri --graph > $OUT/${TAG}_op_bench.jsonl 2> $OUT/${TAG}_op_bench.err
timeout 200 python tools/tc32_bench.py layers slab net fmt=1 > $OUT/${TAG}_tc32h_layers.txt 2>&1
timeout 300 env DINVK_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --lean > $OUT/${TAG}_bench_under_ncu.log 2>&1
if [ -n "$KERNELS" ]; then
  timeout 300 env DINVK_BENCH_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$KERNELS" -c 4 \
      -o $OUT/${TAG}_full python bench.py --steps 2 --warmup 1 --lean > $OUT/${TAG}_ncu_full.log 2>&1
fi
ls -la $OUT | tail -14
