#!/bin/bash
# round-2 GPU call 20: persistent pipelined gn_apply_up; fp8 corrections on noise_embedding.3 — parity subset, timeline, A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or golden or range or every_pixel or one_step or configured or exact_split" > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest20.log
tail -6 gpurun_out/r02_pytest20.log | cut -c1-250
DD_OUT=gpurun_out/r02_timeline_loop_20.json timeout 300 python profiles/timeline_probe.py 2>&1 | tail -13 | tee gpurun_out/r02_timeline_loop_20.log
DD_ENGINE_LIB=$PWD/diffusiondepth_b200/libddengine_probes.so DD_F8_NE3=0 timeout 300 python profiles/timeline_probe.py 2>&1 | tail -13 | tee gpurun_out/r02_timeline_loop_20_ne3_3pass.log
cat gpurun_out/parity_margins.json | head -60
