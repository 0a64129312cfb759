#!/bin/bash
# H-sharded Middlebury bench (configs[4]) on one box: gpurun --gpus 8 -- 'bash benchmarks/hshard_scale.sh 8 4 2'
set -u
mkdir -p gpurun_out
port=29520
for N in "$@"; do
  for T in nccl p2p; do
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus "$N" --hshard --hshard-transport $T --config middlebury_1536x2048 \
      --steps 10 --warmup 3 > gpurun_out/hshard_${T}_${N}gpu.json 2> gpurun_out/hshard_${T}_${N}gpu.err
    echo "hshard $T N=$N rc=$?"; grep '^{' gpurun_out/hshard_${T}_${N}gpu.json | head -c 400; echo
  done
done
