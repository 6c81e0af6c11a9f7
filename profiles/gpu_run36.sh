#!/bin/bash
# round-2 GPU call 36: ncu --set full rows of the producer kernels of the final binary (16-warp convgen epilogue, window attention on tcgen05)
set -x
mkdir -p gpurun_out
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_umma_kernel|ln_split_kernel|patch_embed" -c 16 -o /tmp/r02s2_prod_s0 -f python profiles/run_forward_once.py > gpurun_out/r02s2_ncu_s0.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_umma_kernel|ln_split_kernel" --launch-skip 44 -c 10 -o /tmp/r02s2_prod_s2 -f python profiles/run_forward_once.py > gpurun_out/r02s2_ncu_s2.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel" --launch-skip 99 -c 19 -o /tmp/r02s2_prod_neck -f python profiles/run_forward_once.py > gpurun_out/r02s2_ncu_neck.log 2>&1
for f in r02s2_prod_s0 r02s2_prod_s2 r02s2_prod_neck; do ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
python profiles/ncu_summary.py gpurun_out/r02s2_prod_s0.raw.csv gpurun_out/r02s2_prod_s2.raw.csv gpurun_out/r02s2_prod_neck.raw.csv > gpurun_out/r02s2_producers_summary.csv
wc -l gpurun_out/r02s2_producers_summary.csv; cut -c1-200 gpurun_out/r02s2_producers_summary.csv | head -50
