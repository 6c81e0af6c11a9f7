#!/bin/bash
# round-2 GPU call 6: row-halo swapped-operand kernel, 16-pixel gn_apply_up segments, coalesced patch embed
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -15 gpurun_out/r02_pytest6.log | cut -c1-300
AB_ONLY=1 timeout 600 python profiles/ab_probe.py 2>&1 | tee gpurun_out/r02_ab_probe6.log
DD_ENGINE_LIB=$PWD/diffusiondepth_b200/libddengine_probes.so DD_SWAPHALO=0 AB_ONLY=1 timeout 600 python profiles/ab_probe.py 2>&1 | sed 's/r2 lib/probes lib, DD_SWAPHALO=0/' | tee -a gpurun_out/r02_ab_probe6.log
AB_TAG=fp8 timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee gpurun_out/r02_forward_split6.log
