#!/bin/bash
# round-2 GPU call 8: window attention on tcgen05 (vs the SIMT kernel through the probes build), float4 patch embed
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "swin or golden or native or dropin or main_py" > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
tail -25 gpurun_out/r02_pytest8.log | cut -c1-250
AB_TAG=umma-attention timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee gpurun_out/r02_forward_split8.log
DD_ENGINE_LIB=$PWD/diffusiondepth_b200/libddengine_probes.so DD_ATTN_SIMT=1 AB_TAG=simt-attention timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee -a gpurun_out/r02_forward_split8.log
DD_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_T1_b.csv python profiles/run_forward_once.py > /dev/null 2>&1
grep -c . gpurun_out/r02_launches_forward_T1_b.csv
