#!/bin/bash
# round 2, call 2 (2 GPUs): tile-set tests through the C ABI + the 2-GPU bench line
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/r02_c2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_tiles_gpu.py -x -q > $OUT/r02_c2_pytest_tiles.log 2>&1; echo "tiles pytest rc=$?"; tail -15 $OUT/r02_c2_pytest_tiles.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $OUT/r02_c2_bench2.json 2> $OUT/r02_c2_bench2.err; echo "bench2 rc=$?"
tail -c 3000 $OUT/r02_c2_bench2.json; tail -5 $OUT/r02_c2_bench2.err
timeout 300 python bench.py --steps 50 --warmup 5 > $OUT/r02_c2_bench1.json 2> $OUT/r02_c2_bench1.err; echo "bench1 rc=$?"
tail -c 6000 $OUT/r02_c2_bench1.json; tail -5 $OUT/r02_c2_bench1.err
