#!/bin/bash
# bash tools/variant_sweep.sh "sfx1 sfx2 ..." -- frame times (tools/frame_sweep.py) of library builds side by side at the bench's size and
# at 512^2 x 64 / 2048^2 x 4, the default build first and last (drift check); then the parity tests on each variant.
set -u
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/variant_sweep.jsonl
for sfx in "" $1 ""; do
  for cfg in "1024 16" "512 64" "2048 4"; do
    set -- $cfg
    MW_LIB_SUFFIX=$sfx MW_SWEEP_N=$1 MW_SWEEP_TILES=$2 timeout 200 python tools/frame_sweep.py 2>>$OUT/variant_sweep.err >> $OUT/variant_sweep.jsonl
  done
done
cat $OUT/variant_sweep.jsonl
for sfx in ${VARIANTS_TO_TEST:-}; do
  echo "parity tests on libmistral_ocean$sfx.so"; MW_LIB_SUFFIX=$sfx timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_tiles_gpu.py -x -q -m gpu 2>&1 | tail -2
done
