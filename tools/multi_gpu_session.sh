#!/bin/bash
# Multi-GPU evidence in ONE gpurun call (charged N x): the configurations BASELINE.json quotes on several GPUs plus strong scaling of cfg2
# with the bit-exactness check of the gathered shards.
#   usage: /usr/local/graft/bin/gpurun --gpus 4 --timeout 900 -- 'bash tools/multi_gpu_session.sh 4 r02'
set -u
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, bench args...
  local name=$1; shift
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" \
      > $OUT/${TAG}_bench_${name}_${N}gpu.json 2> $OUT/${TAG}_bench_${name}_${N}gpu.err
  cut -c1-220 $OUT/${TAG}_bench_${name}_${N}gpu.json
}
run cfg2_strong --scaling strong --verify-shards --steps 10 --warmup 3 --lean
run cfg2_weak --steps 10 --warmup 3 --lean
run cfg5 --config cfg5 --steps 5 --warmup 3 --lean
run cfg4 --config cfg4 --steps 5 --warmup 3 --lean
run cfg3 --config cfg3 --steps 5 --warmup 3 --lean
