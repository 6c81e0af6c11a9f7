#!/bin/bash
# round-2 GPU call 3: parity margins (stdout), forward split, GEMM table pairs vs single, ncu --set full on the loop convs
# (fp8 mode) and the producer kernels; only CSV summaries are kept (the .ncu-rep files exceed the 64 MiB return limit)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "golden or exact_split or full_resolution or configured or dropin or main_py" > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
grep "parity\]\|passed\|failed\|Error" gpurun_out/r02_pytest3.log | cut -c1-400
AB_TAG=fp8 timeout 600 python profiles/forward_split.py 2>&1 | tail -2 | tee gpurun_out/r02_forward_split.log
AB_TAG=exact AB_FP8=0 timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee -a gpurun_out/r02_forward_split.log
DD_ENGINE_LIB=$PWD/diffusiondepth_b200/libddengine_probes.so DD_GENPAIR=0 AB_TAG=fp8-genpair-off timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee -a gpurun_out/r02_forward_split.log
timeout 600 python profiles/swin_gemm_table.py > gpurun_out/r02_gemm_table_pair.log 2>&1; tail -25 gpurun_out/r02_gemm_table_pair.log
DD_ENGINE_LIB=$PWD/diffusiondepth_b200/libddengine_probes.so DD_GENPAIR=0 timeout 600 python profiles/swin_gemm_table.py > gpurun_out/r02_gemm_table_single.log 2>&1; tail -25 gpurun_out/r02_gemm_table_single.log
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_swap_kernel|gn_apply|gn_relu_ddim|gn_finalize|decoder_kernel" -c 16 -o /tmp/r02_loop -f python profiles/run_loop_once.py > gpurun_out/r02_ncu_loop.log 2>&1
ncu -i /tmp/r02_loop.ncu-rep --page raw --csv > gpurun_out/r02_loop.raw.csv 2>/dev/null
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel|patch_embed" -c 16 -o /tmp/r02_prod_s0 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s0.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel" --launch-skip 44 -c 10 -o /tmp/r02_prod_s2 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s2.log 2>&1
for f in r02_prod_s0 r02_prod_s2; do ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out /tmp/*.ncu-rep
