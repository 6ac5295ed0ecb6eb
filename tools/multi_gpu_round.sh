#!/bin/bash
# bash tools/multi_gpu_round.sh W [probe cfgs...] -- W GPUs: push-engine probe (optional), the world-W tile check, the bench line at W
set -u
N=${1:-4}; shift
OUT=gpurun_out; mkdir -p $OUT
if [ $# -gt 0 ]; then MW_PUSH_CHECK="tma" bash tools/push_call.sh $N "$@"; cp $OUT/push_probe_$N.jsonl $OUT/r02_push_probe_${N}gpu.jsonl; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/r02_f_bench$N.json 2> $OUT/r02_f_bench$N.err; echo "bench$N rc=$?"
tail -3 $OUT/r02_f_bench$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_f_bench$N.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','multi_gpu')}, indent=1)[:3000])
PY
