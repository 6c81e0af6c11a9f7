#!/bin/bash
# round-2 GPU call 35: final tree — full GPU suite, smoke, bench line (C3), in-situ timelines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest35.log
tail -4 gpurun_out/r02_pytest35.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_s2.json 2> gpurun_out/r02_bench_c3_s2.err; cut -c1-300 gpurun_out/r02_bench_c3_s2.json; tail -3 gpurun_out/r02_bench_c3_s2.err
DD_OUT=gpurun_out/r02_timeline_loop_s2.json timeout 300 python profiles/timeline_probe.py 2>&1 | grep -v "_warn\|UserWarning" | tail -13 | cut -c1-170 | tee gpurun_out/r02_timeline_loop_s2.log
DD_FULL=1 DD_DUMP=gpurun_out/r02_forward_launches_insitu_s2.txt DD_OUT=gpurun_out/r02_timeline_forward_s2.json timeout 400 python profiles/timeline_probe.py 2>&1 | grep -v "_warn\|UserWarning" | tail -28 | cut -c1-170 | tee gpurun_out/r02_timeline_forward_s2.log
