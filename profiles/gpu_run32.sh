#!/bin/bash
# round-2 GPU call 32: gn_apply_up — 4 blocks per SM (64 registers, small spills) and a8 from the fp16 value (timing A/B, alternating)
set -x
mkdir -p gpurun_out
for i in 1 2; do
  for v in libddengine.so libddengine_LB4.so libddengine_H2A8.so; do
  DD_ENGINE_LIB=$PWD/diffusiondepth_b200/$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|gn_apply_up\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_32_${v}_$i.log
  done
done
