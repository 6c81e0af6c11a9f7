#!/bin/bash
# round-2 GPU call 28: 64->256 fp8 kernel with 16 epilogue warps on 16-column chunks vs 8 warps on 32 (same box, alternating) + parity subset
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or golden or range or one_step or every_pixel or configured" > gpurun_out/r02_pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest28.log
tail -4 gpurun_out/r02_pytest28.log | cut -c1-250
for i in 1 2 3; do
  for v in libddengine.so libddengine_epi8.so; do
  DD_ENGINE_LIB=$PWD/diffusiondepth_b200/$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|halo_kernel<64\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_28_${v}_$i.log
  done
done
