#!/bin/bash
# Runs on the GPU box (under gpurun): launch list + full capture of the two frame kernels + Gerstner, for profiles/.
# Usage: tools/profile_round.sh r01
set -u
R=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# (1) every launch of the bench command with its device time (shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${R}_launches_bench.log 2>&1
# (2) full capture of the frame kernels in the batched configuration (one launch covers all 16 tiles)
MW_GROUP_TILES=16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 6 -c 2 \
    -o $OUT/${R}_frame_batched python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${R}_ncu_batched.log 2>&1
# (3) same, default scheduling (one tile per launch, L2-resident intermediate)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cols_extract|k_spectrum_rows" -s 40 -c 2 \
    -o $OUT/${R}_frame_grouped python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${R}_ncu_grouped.log 2>&1
# (4) Gerstner 32 waves x 1M vertices
timeout 300 ncu --set full --clock-control none -k regex:k_gerstner -c 1 -o $OUT/${R}_gerstner python tools/bench_extra.py --only gerstner > $OUT/${R}_ncu_gerstner.log 2>&1
# (4b) OceanRenderer path: Ocean Demo scene (1024^2 maps), 16 oceans per call
timeout 300 ncu --set full --clock-control none -k regex:"k_r_rows|k_r_cols|k_r_maps" -s 6 -c 3 -o $OUT/${R}_renderer python tools/bench_extra.py --only renderer16 > $OUT/${R}_ncu_renderer.log 2>&1
# (4c) where the time of the two frame kernels goes: phases switched off one at a time, one launch for all 16 tiles
MW_GROUP_TILES=16 timeout 200 python tools/phase_timing.py > $OUT/${R}_phases.txt 2>&1
# (5) the bench line itself + the extra configs (no profiler attached)
timeout 300 python bench.py > $OUT/${R}_bench.json 2> $OUT/${R}_bench.err
timeout 300 python tools/bench_extra.py > $OUT/${R}_extra.json 2> $OUT/${R}_extra.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${R}_smi.csv
ls -la $OUT | tail -20
