#!/bin/bash
# round-2 GPU call 2: full GPU suite (fp8 corrections, pairs for 256->256, graphs for producers, Vis path, drop-in, parity
# pins), loop timing fp8 vs exact, ncu --set full on producer kernels, bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_box2.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -40 gpurun_out/r02_pytest2.log
AB_ONLY=1,2 timeout 600 python profiles/ab_probe.py > gpurun_out/r02_ab_probe2.log 2>&1
cat gpurun_out/r02_ab_probe2.log
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel|patch_embed" -c 16 -o gpurun_out/r02_prod_s0 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s0.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel" --launch-skip 44 -c 10 -o gpurun_out/r02_prod_s2 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s2.log 2>&1
for f in r02_prod_s0 r02_prod_s2; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
cat gpurun_out/r02_bench2.json; tail -5 gpurun_out/r02_bench2.err
