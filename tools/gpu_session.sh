#!/bin/bash
# One gpurun call that produces everything a round needs from the B200: the GPU test suite, smoke(), the bench line, the
# operator / training micro-benchmarks, an ncu launch list of one bench step and `--set full` captures of the named kernels.
#   usage (from the repo root, in the authoring container):
#     /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_session.sh r02a "conv_tc_halo|sp_row_fused"'
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/ (tools/ncu_extract.py condenses ncu files).
set -u
TAG=${1:-sess}
KERNELS=${2:-}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; tail -2 $OUT/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-200 $OUT/${TAG}_bench.json
timeout 240 python tools/op_bench.py mri tomo blur mcmri --graph > $OUT/${TAG}_op_bench.jsonl 2> $OUT/${TAG}_op_bench.err
timeout 200 python tools/op_bench.py train > $OUT/${TAG}_op_train.jsonl 2> $OUT/${TAG}_op_train.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
if [ -n "$KERNELS" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -c 6 -o $OUT/${TAG}_full \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
fi
ls -la $OUT | tail -12
