#!/bin/bash
# Runs on the GPU box (under gpurun): everything profiles/ is made from.  Usage: tools/profile_round.sh r02
# Then, here: python tools/summarize_profiles.py r02
set -u
R=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# (1) every launch of the bench command with its device time (shares, not absolutes)
timeout 300 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $OUT/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --e2e-steps 1 > $OUT/${R}_launches_bench.log 2>&1
# (2) full capture of the two frame kernels in the timed scheduling (one tile per launch; caches left alone so that the
#     intermediate is L2-resident as it is in the timed region)
timeout 400 $NCU --set full --cache-control none --import-source on -k regex:"k_cols_|k_spectrum_rows" -s 60 -c 2 \
    -o $OUT/${R}_frame_grouped python tools/traffic_frame.py > $OUT/${R}_ncu_grouped.log 2>&1
# (3) one launch for all 16 tiles (steady state over many waves)
MW_GROUP_TILES=16 timeout 400 $NCU --set full --import-source on -k regex:"k_cols_|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/${R}_frame_batched python tools/traffic_frame.py > $OUT/${R}_ncu_batched.log 2>&1
# (4) the other sizes the configs name (no embedded sources: gpurun brings back at most 64 MiB)
MW_TR_N=2048 MW_TR_TILES=1 timeout 400 $NCU --set full -k regex:"k_cols_|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/${R}_frame_2048 python tools/traffic_frame.py > $OUT/${R}_ncu_2048.log 2>&1
MW_TR_N=256 MW_TR_TILES=256 MW_GROUP_TILES=256 timeout 400 $NCU --set full -k regex:"k_cols_|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/${R}_frame_256 python tools/traffic_frame.py > $OUT/${R}_ncu_256.log 2>&1
# (5) Gerstner 32 waves x 1M vertices; OceanRenderer path (Ocean Demo scene, 16 oceans per call)
timeout 300 $NCU --set full -k regex:k_gerstner -c 1 -o $OUT/${R}_gerstner python tools/bench_extra.py --only gerstner > $OUT/${R}_ncu_gerstner.log 2>&1
timeout 300 $NCU --set full -k regex:"k_r_rows|k_r_cols|k_r_maps" -s 6 -c 3 -o $OUT/${R}_renderer python tools/bench_extra.py --only renderer16 > $OUT/${R}_ncu_renderer.log 2>&1
# (6) DRAM bytes of ONE whole 16-tile frame, every launch, in the timed scheduling and in the one-launch-per-kernel mode
for mode in grouped batched; do
  if [ $mode = batched ]; then export MW_GROUP_TILES=16; SK="-s 9 -c 3"; else unset MW_GROUP_TILES; SK="-s 99 -c 33"; fi
  timeout 400 $NCU --replay-mode application --cache-control none \
     --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
     -k regex:"k_cols_|k_spectrum_rows|k_phase_table" $SK --csv --log-file $OUT/${R}_traffic_$mode.csv \
     python tools/traffic_frame.py > $OUT/${R}_traffic_$mode.log 2>&1
  python tools/summarize_traffic.py $OUT/${R}_traffic_$mode.csv > $OUT/${R}_traffic_$mode.json
done
unset MW_GROUP_TILES
# (7) the bench line itself + the OceanRenderer figures (no profiler attached)
timeout 400 python bench.py > $OUT/${R}_bench.json 2> $OUT/${R}_bench.err
timeout 300 python tools/bench_extra.py --only renderer > $OUT/${R}_extra.json 2> $OUT/${R}_extra.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${R}_smi.csv
du -sh $OUT; ls -la $OUT | tail -25
