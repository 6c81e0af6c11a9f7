#!/bin/bash
# round-2 GPU call 38: the `*Vis` heads against goldens from the REAL reference's Vis heads (pred_inter step by step)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "vis" > gpurun_out/r02_pytest38.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest38.log
grep -n "parity\]\|\[vis\]\|passed\|failed\|Error\|assert" gpurun_out/r02_pytest38.log | cut -c1-300 | tail -20
