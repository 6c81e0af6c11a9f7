#!/bin/bash
# round-2 GPU call 19: gn_apply_up with bulk-copied conv outputs + leaner arithmetic — parity subset + in-situ timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or golden or range or every_pixel or one_step" > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest19.log
tail -6 gpurun_out/r02_pytest19.log | cut -c1-250
DD_OUT=gpurun_out/r02_timeline_loop_19.json timeout 300 python profiles/timeline_probe.py 2>&1 | tail -14 | tee gpurun_out/r02_timeline_loop_19.log
