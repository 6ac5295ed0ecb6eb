#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/r02_c7_sweep.jsonl
for v in "" _w4 _p32 _p32w4 _p32w4r2; do
  echo "== variant [$v]"
  MW_LIB_SUFFIX=$v timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -k "1024" 2>&1 | tail -1
  MW_LIB_SUFFIX=$v timeout 200 python tools/frame_sweep.py >> $OUT/r02_c7_sweep.jsonl 2>>$OUT/r02_c7_sweep.err
  MW_LIB_SUFFIX=$v MW_GROUP_TILES=16 timeout 200 python tools/frame_sweep.py >> $OUT/r02_c7_sweep.jsonl 2>>$OUT/r02_c7_sweep.err
done
MW_GROUP_TILES=2 timeout 200 python tools/frame_sweep.py >> $OUT/r02_c7_sweep.jsonl 2>>$OUT/r02_c7_sweep.err
MW_LIB_SUFFIX=_p32w4 MW_GROUP_TILES=2 timeout 200 python tools/frame_sweep.py >> $OUT/r02_c7_sweep.jsonl 2>>$OUT/r02_c7_sweep.err
cat $OUT/r02_c7_sweep.jsonl; tail -3 $OUT/r02_c7_sweep.err
