#!/bin/bash
# round-2 GPU call 26: gn_apply_up one quad per block vs four (same box, alternating); neck alt N tile; parity subset
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "denoiser_operator or loop_and_decode or neck_and_fpn or golden or range or one_step" > gpurun_out/r02_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest26.log
tail -4 gpurun_out/r02_pytest26.log | cut -c1-250
P=$PWD/diffusiondepth_b200/libddengine_probes.so
for i in 1 2; do
  for v in 1 4; do
  DD_ENGINE_LIB=$P DD_UP_QPB=$v timeout 300 python profiles/timeline_probe.py 2>&1 | grep "halo_kernel<256\|gn_apply_up\|kernels in one" | cut -c1-170 | tee gpurun_out/r02_timeline_26_qpb${v}_$i.log
  done
done
DD_FULL=1 DD_DUMP=gpurun_out/r02_forward_launches_insitu_26.txt timeout 400 python profiles/timeline_probe.py 2>&1 | grep "kernels in one\|convgen" | cut -c1-170
sed -n 180,198p gpurun_out/r02_forward_launches_insitu_26.txt
