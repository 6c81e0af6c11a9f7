#!/bin/bash
# round-2 GPU call 23: gn_apply_up quad kernel, 4 vs 8 channels per thread (same box, alternating) + parity subset
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or golden or range or one_step" > gpurun_out/r02_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest23.log
tail -4 gpurun_out/r02_pytest23.log | cut -c1-250
P=$PWD/diffusiondepth_b200/libddengine_probes.so
for i in 1 2; do
  for v in 4 8; do
  DD_ENGINE_LIB=$P DD_UP_VEC=$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|gn_apply_up\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_23_v${v}_$i.log
  done
done
