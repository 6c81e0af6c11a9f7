#!/bin/bash
# round-2 GPU call 10: attention with 3 CTAs / SM (P over Q|K, O over S), decoder v2, patch embed; full suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest10.log
tail -8 gpurun_out/r02_pytest10.log | cut -c1-250
AB_TAG=run10 timeout 600 python profiles/forward_split.py 2>&1 | tail -1 | tee gpurun_out/r02_forward_split10.log
DD_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_T1_c.csv python profiles/run_forward_once.py > /dev/null 2>&1
python - <<'PY'
import csv,re
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_forward_T1_c.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name')
import collections
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki])[:60]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{n:4d} {t/1e3:9.1f} us {k}")
PY
