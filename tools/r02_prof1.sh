#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
# default scheduling (one tile per launch, intermediate L2-resident): one launch of each frame kernel, caches left alone
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 60 -c 2 \
    -o $OUT/r02_frame_grouped python tools/traffic_frame.py > $OUT/r02_ncu_grouped.log 2>&1
# one launch for all 16 tiles (steady state, many waves)
MW_GROUP_TILES=16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/r02_frame_batched python tools/traffic_frame.py > $OUT/r02_ncu_batched.log 2>&1
# 2048^2 single tile
MW_TR_N=2048 MW_TR_TILES=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/r02_frame_2048 python tools/traffic_frame.py > $OUT/r02_ncu_2048.log 2>&1
# 256^2 x 256
MW_TR_N=256 MW_TR_TILES=256 MW_GROUP_TILES=256 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/r02_frame_256 python tools/traffic_frame.py > $OUT/r02_ncu_256.log 2>&1
ls -la $OUT/*.ncu-rep
