#!/bin/bash
# round-2 GPU call 17 (session 2): verification of the restored tree — full GPU suite, smoke, bench line, in-situ layer timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest17.log
tail -6 gpurun_out/r02_pytest17.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_17.json 2> gpurun_out/r02_bench_c3_17.err; cut -c1-400 gpurun_out/r02_bench_c3_17.json; tail -3 gpurun_out/r02_bench_c3_17.err
timeout 300 python profiles/loop_layers.py 2>&1 | tail -9 | tee gpurun_out/r02_loop_layers_17.log
timeout 300 python profiles/forward_split.py 2>&1 | tail -6 | tee gpurun_out/r02_forward_split_17.log
