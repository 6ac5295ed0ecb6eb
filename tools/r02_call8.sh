#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02_c8_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/r02_c8_pytest.log
: > $OUT/r02_c8_sweep.jsonl
for v in "" _c88 _c80; do
  MW_LIB_SUFFIX=$v MW_SWEEP_N=256 MW_SWEEP_TILES=256 timeout 200 python tools/frame_sweep.py >> $OUT/r02_c8_sweep.jsonl 2>>$OUT/r02_c8_sweep.err
done
MW_SWEEP_N=64 MW_SWEEP_TILES=1 timeout 100 python tools/frame_sweep.py >> $OUT/r02_c8_sweep.jsonl 2>>$OUT/r02_c8_sweep.err
MW_INLINE_PHASE=0 MW_SWEEP_N=64 MW_SWEEP_TILES=1 timeout 100 python tools/frame_sweep.py >> $OUT/r02_c8_sweep.jsonl 2>>$OUT/r02_c8_sweep.err
cat $OUT/r02_c8_sweep.jsonl; tail -3 $OUT/r02_c8_sweep.err
bash tools/sanitize.sh > $OUT/r02_sanitizer.txt 2>&1; tail -30 $OUT/r02_sanitizer.txt
