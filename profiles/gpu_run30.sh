#!/bin/bash
# round-2 GPU call 30: what paces the 64->256 fp8 kernel — ring depths (3 strip units + 2 weight stages) and a no-store epilogue (timing only)
set -x
mkdir -p gpurun_out
for i in 1 2; do
  for v in libddengine.so libddengine_A3B2.so libddengine_NOSTORE.so; do
  DD_ENGINE_LIB=$PWD/diffusiondepth_b200/$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|halo_kernel<64\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_30_${v}_$i.log
  done
done
