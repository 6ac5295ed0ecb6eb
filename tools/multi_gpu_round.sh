#!/bin/bash
# N GPUs (argument): the world-N tile-set test + the bench line at N
set -u
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02_c5_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_tiles_gpu.py -x -q > $OUT/r02_c5_pytest_tiles_$N.log 2>&1; echo "tiles pytest rc=$?"; tail -5 $OUT/r02_c5_pytest_tiles_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > $OUT/r02_c5_bench$N.json 2> $OUT/r02_c5_bench$N.err; echo "bench$N rc=$?"
tail -3 $OUT/r02_c5_bench$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_c5_bench$N.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','multi_gpu')}, indent=1)[:3500])
PY
