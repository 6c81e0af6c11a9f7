#!/bin/bash
# round-2 GPU call 21: same-box A/B of fp8 corrections on noise_embedding.3 (alternating, per-launch medians)
set -x
mkdir -p gpurun_out
P=$PWD/diffusiondepth_b200/libddengine_probes.so
for i in 1 2; do
  DD_ENGINE_LIB=$P DD_F8_NE3=1 timeout 300 python profiles/timeline_probe.py 2>&1 | grep -v "^  _warn\|UserWarning" | tail -12 | cut -c1-170 | tee gpurun_out/r02_timeline_21_f8ne3_$i.log
  DD_ENGINE_LIB=$P DD_F8_NE3=0 timeout 300 python profiles/timeline_probe.py 2>&1 | grep -v "^  _warn\|UserWarning" | tail -12 | cut -c1-170 | tee gpurun_out/r02_timeline_21_3pass_$i.log
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
