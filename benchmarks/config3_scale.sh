#!/bin/bash
# BASELINE.json configs[2]: SceneFlow 544x960, maxdisp 192, batch 64 split 64/N pairs per GPU, N in the arguments.
#   gpurun --gpus 8 -- 'bash benchmarks/config3_scale.sh 1 2 4 8'
set -u
mkdir -p gpurun_out
port=29540
for N in "$@"; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --config sceneflow_544x960 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --latency-steps 0 \
      > gpurun_out/config3_${N}gpu.json 2> gpurun_out/config3_${N}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$N" --config sceneflow_544x960 --batch 64 --steps 5 --warmup 3 --latency-steps 0 \
      > gpurun_out/config3_${N}gpu.json 2> gpurun_out/config3_${N}gpu.err
  fi
  echo "config3 N=$N rc=$?"
  grep '^{' gpurun_out/config3_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])"
done
