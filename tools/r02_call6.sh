#!/bin/bash
set -u
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gather_probe.py 2>>$OUT/r02_c6_probe.err | grep '^{' >> $OUT/r02_c6_probe_$N.jsonl; }
: > $OUT/r02_c6_probe_$N.jsonl
run
MW_TILES_PUSH_LANES=1 run
MW_TILES_PUSH_LANES=2 run
CUDA_DEVICE_MAX_CONNECTIONS=32 run
CUDA_DEVICE_MAX_CONNECTIONS=32 MW_TILES_PUSH_LANES=1 run
MW_PROBE_TILES=4 run
cat $OUT/r02_c6_probe_$N.jsonl; tail -3 $OUT/r02_c6_probe.err
