#!/bin/bash
# round-2 GPU call 14: MPViT tests again (cached-mirror pollution fixed), MPViT forward timing at KITTI size, launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mpvit" > gpurun_out/r02_pytest14a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest14a.log
grep -n "mpvit stages\|parity\]\|passed\|failed" gpurun_out/r02_pytest14a.log | cut -c1-300
timeout 600 python profiles/mpvit_forward_time.py > gpurun_out/r02_mpvit_forward.json 2> gpurun_out/r02_mpvit_forward.err; cat gpurun_out/r02_mpvit_forward.json; tail -5 gpurun_out/r02_mpvit_forward.err
DD_B=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mpvit_B1.csv python profiles/mpvit_forward_time.py > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_mpvit_B1.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name')
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    n=re.sub(r'>\(.*','>',r[ki])[:80]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[-1]); tot+=float(r[-1])
print('total us', tot/1e3)
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{n:5d} {t/1e3:10.1f} us {k}")
PY
