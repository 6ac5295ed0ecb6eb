set -u
OUT=gpurun_out; : > $OUT/env_sweep.jsonl
run() { env "$@" MW_SWEEP_N=1024 MW_SWEEP_TILES=16 timeout 100 python tools/frame_sweep.py 2>>$OUT/env_sweep.err >> $OUT/env_sweep.jsonl; }
run MW_X=0
run MW_ROWS_MINB=4
run MW_PDL=2
run MW_SLOTS=3
run MW_GROUP_TILES=2
run MW_ROWS_MINB=2
run MW_X=0
cut -c1-200 $OUT/env_sweep.jsonl
