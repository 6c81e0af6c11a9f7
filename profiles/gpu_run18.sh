#!/bin/bash
# round-2 GPU call 18: in-situ kernel timeline (CUPTI) of the loop graph and of the full forward
set -x
mkdir -p gpurun_out
DD_OUT=gpurun_out/r02_timeline_loop.json timeout 300 python profiles/timeline_probe.py 2>&1 | tail -25 | tee gpurun_out/r02_timeline_loop.log
DD_FULL=1 DD_OUT=gpurun_out/r02_timeline_forward.json timeout 400 python profiles/timeline_probe.py 2>&1 | tail -45 | tee gpurun_out/r02_timeline_forward.log
