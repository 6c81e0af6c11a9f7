#!/bin/bash
# round-2 GPU call 4: full suite with parity margins on stdout; loop / forward timing with the 64-channel fp8 layout and
# 64-channel producer GEMM chunks
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
grep "parity\]\|passed\|failed\|Error\|error" gpurun_out/r02_pytest4.log | cut -c1-300 | head -80
AB_ONLY=1,2 timeout 600 python profiles/ab_probe.py 2>&1 | tee gpurun_out/r02_ab_probe4.log
AB_TAG=fp8 timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee gpurun_out/r02_forward_split4.log
timeout 600 python profiles/swin_gemm_table.py > gpurun_out/r02_gemm_table4.log 2>&1; tail -22 gpurun_out/r02_gemm_table4.log
