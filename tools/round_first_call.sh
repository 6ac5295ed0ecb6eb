#!/bin/bash
# First GPU call of a round (run under gpurun from the repo root, after tools/build_variants.sh here):
#   tests, the bench line, the occupancy sweep, then the ncu summaries.  Usage: tools/round_first_call.sh r02
set -u
R=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/${R}_pytest.txt
timeout 300 python bench.py > $OUT/${R}_bench.json 2> $OUT/${R}_bench.err; tail -c 600 $OUT/${R}_bench.json
for v in "" _occ; do
  [ -f mistral-water_b200/lib/libmistral_ocean$v.so ] || continue
  MW_LIB_SUFFIX=$v timeout 200 python tools/occ_sweep.py 2>&1 | tail -1 | tee -a $OUT/${R}_occ_sweep.jsonl
done
MW_LIB_SUFFIX=_occ timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "full_sizes or golden or config2 or properties" 2>&1 | tail -2 | tee $OUT/${R}_pytest_occ.txt
# tile-group size / launch options of the bench workload (one JSON line each)
for v in "MW_PDL=1" "MW_GROUP_TILES=2" "MW_GROUP_TILES=2 MW_SLOTS=3" "MW_GROUP_TILES=4"; do
  env $v timeout 120 python tools/pdl_sweep.py 2>&1 | tail -1 | tee -a $OUT/${R}_sched_sweep.jsonl
done
bash tools/profile_round.sh $R
