#!/bin/bash
# round-2 GPU call 31 (8 GPUs): weak scaling of the batch shard with the final binary of session 2, C3 (= BASELINE config 4 at N = 8)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --workload C3 --no-cpu-baseline > gpurun_out/r02_scale8_C3_s2.json 2> gpurun_out/r02_scale8_C3_s2.err
cat gpurun_out/r02_scale8_C3_s2.json; tail -3 gpurun_out/r02_scale8_C3_s2.err
