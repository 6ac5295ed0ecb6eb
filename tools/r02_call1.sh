#!/bin/bash
# round 2, first GPU call: baseline check + build-switch sweep + DRAM traffic of the timed scheduling
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r02_c1_smi.csv
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/r02_c1_pytest.log 2>&1; echo "pytest rc=$?"
: > $OUT/r02_c1_sweep.jsonl
for v in "" _w4 _tma _w4tma; do
  echo "== variant [$v]"
  MW_LIB_SUFFIX=$v timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -k "1024 or 64" 2>&1 | tail -2
  MW_LIB_SUFFIX=$v timeout 200 python tools/pdl_sweep.py >> $OUT/r02_c1_sweep.jsonl 2>$OUT/r02_c1_sweep_$v.err
  MW_LIB_SUFFIX=$v MW_GROUP_TILES=16 timeout 200 python tools/pdl_sweep.py >> $OUT/r02_c1_sweep.jsonl 2>>$OUT/r02_c1_sweep_$v.err
done
: > $OUT/r02_c1_occ.jsonl
for v in "" _occ; do MW_LIB_SUFFIX=$v timeout 300 python tools/occ_sweep.py >> $OUT/r02_c1_occ.jsonl 2>$OUT/r02_c1_occ_$v.err; done
# DRAM traffic of one whole 16-tile frame in the timed scheduling (no cache flush between kernels, app replay)
timeout 400 ncu --replay-mode application --cache-control none --clock-control none \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
   -k regex:"k_cols_extract|k_spectrum_rows|k_phase_table" -s 99 -c 33 --csv --log-file $OUT/r02_traffic_grouped.csv \
   python tools/traffic_frame.py > $OUT/r02_traffic_grouped.log 2>&1
python tools/summarize_traffic.py $OUT/r02_traffic_grouped.csv > $OUT/r02_traffic_grouped.json
MW_GROUP_TILES=16 timeout 400 ncu --replay-mode application --cache-control none --clock-control none \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
   -k regex:"k_cols_extract|k_spectrum_rows|k_phase_table" -s 9 -c 3 --csv --log-file $OUT/r02_traffic_batched.csv \
   python tools/traffic_frame.py > $OUT/r02_traffic_batched.log 2>&1
python tools/summarize_traffic.py $OUT/r02_traffic_batched.csv > $OUT/r02_traffic_batched.json
cat $OUT/r02_c1_sweep.jsonl $OUT/r02_c1_occ.jsonl; cat $OUT/r02_traffic_grouped.json | head -40
