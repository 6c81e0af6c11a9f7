#!/bin/bash
# round-2 GPU call 7: committed evidence — launch lists, ncu --set full summaries (loop + producers), benches C3 / C2 / C5
set -x
mkdir -p gpurun_out
DD_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_loop_T2.csv python profiles/run_loop_once.py > /dev/null 2>&1
DD_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_T1.csv python profiles/run_forward_once.py > /dev/null 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_swap_kernel|gn_apply|gn_relu_ddim|decoder_kernel" -c 12 -o /tmp/r02_loop -f python profiles/run_loop_once.py > gpurun_out/r02_ncu_loop.log 2>&1
ncu -i /tmp/r02_loop.ncu-rep --page raw --csv > gpurun_out/r02_loop.raw.csv 2>/dev/null
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel|patch_embed" -c 16 -o /tmp/r02_prod_s0 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s0.log 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"convgen_umma_kernel|window_attention_kernel|ln_split_kernel" --launch-skip 44 -c 10 -o /tmp/r02_prod_s2 -f python profiles/run_forward_once.py > gpurun_out/r02_ncu_s2.log 2>&1
for f in r02_prod_s0 r02_prod_s2; do ncu -i /tmp/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; cat gpurun_out/r02_bench_c3.json; tail -3 gpurun_out/r02_bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --exact --no-cpu-baseline > gpurun_out/r02_bench_c3_exact.json 2>> gpurun_out/r02_bench_c3.err; cat gpurun_out/r02_bench_c3_exact.json
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; cat gpurun_out/r02_bench_c2.json; tail -3 gpurun_out/r02_bench_c2.err
timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err; cat gpurun_out/r02_bench_c5.json; tail -3 gpurun_out/r02_bench_c5.err
ls -la gpurun_out | head -40
