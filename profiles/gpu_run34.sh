#!/bin/bash
# round-2 GPU call 34: gn_apply_up walking images / quad rows backwards (L2: producer's tail, consumer's head) vs forwards, alternating; parity subset
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or odd or one_step" > gpurun_out/r02_pytest34.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest34.log
tail -3 gpurun_out/r02_pytest34.log | cut -c1-200
for i in 1 2 3; do
  for v in libddengine.so libddengine_FWD.so; do
  DD_ENGINE_LIB=$PWD/diffusiondepth_b200/$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|gn_apply_up\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_34_${v}_$i.log
  done
done
