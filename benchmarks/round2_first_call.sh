#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash benchmarks/round2_first_call.sh'            (1 GPU part)
#   gpurun --gpus 2 --timeout 900 -- 'bash benchmarks/round2_first_call.sh 2'  (H-sharded bench at 2 GPUs)
set -u
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  # 1. the regular parity suite must still be green (host graph was generalised to N cva stages)
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
  # 2. opt-in tests: H-sharded kernel sequence (virtual ranks on one device) and the stage-count variants
  DCA_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_hshard.py tests/test_gpu_variants.py tests/test_gpu_composites.py -q \
      > gpurun_out/r2_unvalidated_tests.log 2>&1; echo "unvalidated tests rc=$?"
  # 3. bench line of the unchanged default route
  timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
  tail -c 600 gpurun_out/r2_unvalidated_tests.log
else
  DCA_TEST_UNVALIDATED=1 timeout 900 python -m pytest tests/test_gpu_hshard.py -q -k two_gpus \
      > gpurun_out/r2_hshard_2gpu_tests.log 2>&1; echo "2-gpu hshard tests rc=$?"; tail -c 400 gpurun_out/r2_hshard_2gpu_tests.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$N" --hshard --config middlebury_1536x2048 --steps 10 --warmup 3 \
      > gpurun_out/r2_hshard_${N}gpu.json 2> gpurun_out/r2_hshard_${N}gpu.err; echo "hshard bench rc=$?"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29512 bench.py --gpus "$N" --hshard --hshard-transport p2p --config middlebury_1536x2048 --steps 10 \
      --warmup 3 > gpurun_out/r2_hshard_p2p_${N}gpu.json 2> gpurun_out/r2_hshard_p2p_${N}gpu.err; echo "hshard p2p bench rc=$?"
  timeout 600 python bench.py --config middlebury_1536x2048 --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r2_middlebury_1gpu.json 2>/dev/null; echo "1-gpu middlebury rc=$?"
fi
