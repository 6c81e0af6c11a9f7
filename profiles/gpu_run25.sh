#!/bin/bash
# round-2 GPU call 25: ncu --set full of the kernels changed in session 2 (gn_apply_up quad kernel, 64->256 with fp8 corrections) + convA
set -x
mkdir -p gpurun_out
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|gn_apply_up" -c 5 -o /tmp/r02b_loop -f python profiles/run_loop_once.py > gpurun_out/r02b_ncu_loop.log 2>&1
ncu -i /tmp/r02b_loop.ncu-rep --page raw --csv > gpurun_out/r02b_loop.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02b_loop.raw.csv > gpurun_out/r02b_loop_summary.csv
cut -c1-400 gpurun_out/r02b_loop_summary.csv
cp /tmp/r02b_loop.ncu-rep gpurun_out/ 2>/dev/null; ls -la gpurun_out/*.ncu-rep
