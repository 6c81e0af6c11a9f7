#!/bin/bash
# round-2 GPU call 12: convgen with arbitrary channel counts (MPViT neck + FPN native); attention at 3 CTAs / SM; full suite; bench
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "neck_and_fpn or mpvit" > gpurun_out/r02_pytest12a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest12a.log
tail -15 gpurun_out/r02_pytest12a.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest12.log
tail -8 gpurun_out/r02_pytest12.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_c.json 2> gpurun_out/r02_bench_c3_c.err; cat gpurun_out/r02_bench_c3_c.json | cut -c1-1200; tail -3 gpurun_out/r02_bench_c3_c.err
