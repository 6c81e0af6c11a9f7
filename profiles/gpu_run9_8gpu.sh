#!/bin/bash
# round-2 GPU call 9 (8 GPUs): weak scaling of the batch shard with the off-stream all-gather: C3 (= BASELINE config 4 at N = 8)
# and C5 (config 5: 8 x 480x640 per GPU, T = 50), per-rank timings in the line
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,power.limit --format=csv | head -10
for W in C3 C5; do
  S=8; [ $W = C5 ] && S=3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps $S --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/r02_scale8_$W.json 2> gpurun_out/r02_scale8_$W.err
  cat gpurun_out/r02_scale8_$W.json; tail -3 gpurun_out/r02_scale8_$W.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --blocking-gather > gpurun_out/r02_scale8_C3_blocking.json 2> gpurun_out/r02_scale8_C3_blocking.err
cat gpurun_out/r02_scale8_C3_blocking.json
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale1_C3.json 2>/dev/null; cat gpurun_out/r02_scale1_C3.json
