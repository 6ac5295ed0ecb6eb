#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02_c3_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/r02_c3_bench2.json 2> $OUT/r02_c3_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c3_bench2.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','multi_gpu')}, indent=1)[:3000])
PY
