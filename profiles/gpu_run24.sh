#!/bin/bash
# round-2 GPU call 24: per-launch dump of the full forward (producers) inside the replayed graphs
set -x
mkdir -p gpurun_out
DD_FULL=1 DD_DUMP=gpurun_out/r02_forward_launches_insitu.txt DD_OUT=gpurun_out/r02_timeline_forward_24.json timeout 400 python profiles/timeline_probe.py 2>&1 | grep -v "_warn\|UserWarning" | tail -30 | cut -c1-170 | tee gpurun_out/r02_timeline_forward_24.log
