#!/bin/bash
# round-2 GPU call 16: ktv_partial over head groups, ILP softmax partials, 2-D tiled apply; timing + launch list
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mpvit" > gpurun_out/r02_pytest16a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest16a.log
grep -n "mpvit stages\|parity\]\|passed\|failed\|Error" gpurun_out/r02_pytest16a.log | cut -c1-300
timeout 600 python profiles/mpvit_forward_time.py > gpurun_out/r02_mpvit_forward.json 2> gpurun_out/r02_mpvit_forward.err; cat gpurun_out/r02_mpvit_forward.json; tail -5 gpurun_out/r02_mpvit_forward.err
DD_NATIVE_ONLY=1 DD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mpvit_B4.csv python profiles/mpvit_forward_time.py > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_mpvit_B4.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name')
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    n=re.sub(r'>\(.*','>',r[ki])[:80]; a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[-1]); tot+=float(r[-1])
print('total us', tot/1e3, '(3 forwards + 2 backbone runs, B=4)')
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print(f"{n:5d} {t/1e3:10.1f} us {k}")
PY
