#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/r02_c4_pytest.log
: > $OUT/r02_c4_sweep.jsonl
timeout 200 python tools/pdl_sweep.py >> $OUT/r02_c4_sweep.jsonl 2>$OUT/r02_c4_sweep.err
MW_GRAPH=0 timeout 200 python tools/pdl_sweep.py >> $OUT/r02_c4_sweep.jsonl 2>>$OUT/r02_c4_sweep.err
cat $OUT/r02_c4_sweep.jsonl
timeout 300 python bench.py --steps 50 --warmup 5 > $OUT/r02_c4_bench1.json 2> $OUT/r02_c4_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c4_bench1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
print({k:(v.get('us_per_frame') or v.get('us')) for k,v in d['configs'].items()})
PY
