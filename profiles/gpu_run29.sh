#!/bin/bash
# round-2 GPU call 29 (session 2 evidence, final binary of the session: 16-warp epilogue on the 64->256 fp8 kernel): full GPU suite, smoke, launch lists, ncu --set full of the loop kernels, bench lines, in-situ timelines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest29.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest29.log
tail -5 gpurun_out/r02_pytest29.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
DD_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_loop_T2_s2.csv python profiles/run_loop_once.py > /dev/null 2>&1
DD_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_T1_s2.csv python profiles/run_forward_once.py > /dev/null 2>&1
DD_STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo_kernel|conv3x3_swap_kernel|gn_apply|gn_relu_ddim|decoder_kernel" -c 12 -o /tmp/r02s2_loop -f python profiles/run_loop_once.py > gpurun_out/r02s2_ncu_loop.log 2>&1
ncu -i /tmp/r02s2_loop.ncu-rep --page raw --csv > gpurun_out/r02s2_loop.raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02s2_loop.raw.csv > gpurun_out/r02s2_loop_summary.csv; wc -l gpurun_out/r02s2_loop_summary.csv
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_s2.json 2> gpurun_out/r02_bench_c3_s2.err; cut -c1-300 gpurun_out/r02_bench_c3_s2.json; tail -3 gpurun_out/r02_bench_c3_s2.err
timeout 600 python bench.py --steps 10 --warmup 3 --exact --no-cpu-baseline > gpurun_out/r02_bench_c3_exact_s2.json 2>> gpurun_out/r02_bench_c3_s2.err; cut -c1-200 gpurun_out/r02_bench_c3_exact_s2.json
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_s2.json 2> gpurun_out/r02_bench_c2_s2.err; cut -c1-200 gpurun_out/r02_bench_c2_s2.json; tail -3 gpurun_out/r02_bench_c2_s2.err
timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_s2.json 2> gpurun_out/r02_bench_c5_s2.err; cut -c1-200 gpurun_out/r02_bench_c5_s2.json; tail -3 gpurun_out/r02_bench_c5_s2.err
DD_OUT=gpurun_out/r02_timeline_loop_s2.json timeout 300 python profiles/timeline_probe.py 2>&1 | grep -v "_warn\|UserWarning" | tail -13 | cut -c1-170 | tee gpurun_out/r02_timeline_loop_s2.log
DD_FULL=1 DD_DUMP=gpurun_out/r02_forward_launches_insitu_s2.txt DD_OUT=gpurun_out/r02_timeline_forward_s2.json timeout 400 python profiles/timeline_probe.py 2>&1 | grep -v "_warn\|UserWarning" | tail -28 | cut -c1-170 | tee gpurun_out/r02_timeline_forward_s2.log
