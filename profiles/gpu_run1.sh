#!/bin/bash
# round-2 GPU call 1: sanity tests on the restructured producers, A/B probes, ncu --set full on producer kernels, bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_box1.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -3 gpurun_out/r02_pytest1.log
timeout 900 python profiles/ab_probe.py > gpurun_out/r02_ab_probe.log 2>&1
cat gpurun_out/r02_ab_probe.log
# producers: stage-0 kernels (first 16 matching launches) and stage-2 (skip 40)
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel|patch_embed" -c 16 -o gpurun_out/r02_prod_s0 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s0.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel" --launch-skip 44 -c 10 -o gpurun_out/r02_prod_s2 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s2.log 2>&1
for f in r02_prod_s0 r02_prod_s2; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
cat gpurun_out/r02_bench1.json
