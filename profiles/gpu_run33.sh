#!/bin/bash
# round-2 GPU call 33: the odd-size trained-like Swin golden from the real reference (70x106: padded patch embed, 18x27 condition under a
# 35x53 latent = per-tap path of the condition-injection kernel, resampling FPN, odd pyramid levels)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "odd" > gpurun_out/r02_pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest33.log
grep -n "parity\]\|passed\|failed\|Error\|assert" gpurun_out/r02_pytest33.log | cut -c1-300 | tail -20
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_margins.json'))
for r in d['rows']: print(r['case'], r['against'][:40], '%.2e %.2e'%(r['max_dz'], r['rms_dz']), r.get('latent_rel'), r.get('cond_rel'))
PY
