#!/bin/bash
# round-2 GPU call 37 (2 GPUs): the torchrun launch of bench.py prints exactly one JSON line on stdout (NCCL banner on stderr)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale2_C3_s2.json 2> gpurun_out/r02_scale2_C3_s2.err
wc -l gpurun_out/r02_scale2_C3_s2.json; cut -c1-200 gpurun_out/r02_scale2_C3_s2.json; grep -c "NCCL version" gpurun_out/r02_scale2_C3_s2.err; tail -2 gpurun_out/r02_scale2_C3_s2.err | cut -c1-200
