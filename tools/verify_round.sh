#!/bin/bash
# What the driver does at round end, on one GPU: the GPU tests, smoke(), the bench line and the reference arm.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $OUT/verify_bench.json 2> $OUT/verify_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/verify_ref.json 2> $OUT/verify_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/verify_bench.json')); r=json.load(open('gpurun_out/verify_ref.json'))
print(round(d['value']/1e9,2), round(d['ms_per_step']*1e3,1), d['roofline']['frac'], round(d['e2e']['value']/1e9,3), d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
print({k:(v.get('us_per_frame') or v.get('us')) for k,v in d['configs'].items()})
print('ref', r['value'], r['cpu_baseline']['cores'], r['config']==d['config'])
PY
