#!/bin/bash
# round-2 GPU call 5: 16-warp / 16-column epilogue of the producer GEMM kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "native or golden or neck or backbone" > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -5 gpurun_out/r02_pytest5.log
AB_TAG=fp8 timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee gpurun_out/r02_forward_split5.log
timeout 600 python profiles/swin_gemm_table.py > gpurun_out/r02_gemm_table5.log 2>&1; tail -22 gpurun_out/r02_gemm_table5.log
