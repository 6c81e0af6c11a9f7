#!/bin/bash
# round-2 GPU call 39: full GPU suite of the final tree (goldens incl. the odd-size Swin and the two Vis cases)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest39.log
tail -4 gpurun_out/r02_pytest39.log | cut -c1-250
