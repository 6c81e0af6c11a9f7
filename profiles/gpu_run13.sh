#!/bin/bash
# round-2 GPU call 13: native MPViT backbone — focused tests first, then the full suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mpvit" > gpurun_out/r02_pytest13a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest13a.log
tail -40 gpurun_out/r02_pytest13a.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest13.log
tail -12 gpurun_out/r02_pytest13.log | cut -c1-300
