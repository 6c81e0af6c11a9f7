#!/bin/bash
# round-2 GPU call 22: quad-based gn_apply_up (source reuse in registers) — parity subset + timeline x2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or golden or range or every_pixel or one_step or configured or exact_split" > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest22.log
tail -6 gpurun_out/r02_pytest22.log | cut -c1-250
for i in 1 2; do
DD_OUT=gpurun_out/r02_timeline_loop_22_$i.json timeout 300 python profiles/timeline_probe.py 2>&1 | grep -v "^  _warn\|UserWarning" | tail -12 | cut -c1-170 | tee gpurun_out/r02_timeline_loop_22_$i.log
done
