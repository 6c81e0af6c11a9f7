#!/bin/bash
# round-2 GPU call 11: attention with 4 CTAs / SM; full suite; bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest11.log
tail -6 gpurun_out/r02_pytest11.log | cut -c1-250
DD_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_T1_d.csv python profiles/run_forward_once.py > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_forward_T1_d.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name')
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki])[:60]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"{n:4d} {t/1e3:9.1f} us {k}")
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_b.json 2> gpurun_out/r02_bench_c3_b.err; cat gpurun_out/r02_bench_c3_b.json | cut -c1-1200; tail -3 gpurun_out/r02_bench_c3_b.err
