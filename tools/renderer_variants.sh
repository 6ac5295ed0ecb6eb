#!/bin/bash
# OceanRenderer path, library variants side by side (built here with MW_LIB_SUFFIX / MW_NVCC_DEFS, see mistral-water_b200/build.py):
#   _rb  per-thread slab loads, k_r_maps once at the end (the round-2 start)      _rt  bulk-copy slab loads only
#   ""   bulk-copy slab loads + k_r_maps per tile group (default)                   _ru3 default + evolve loop unrolled x3
set -u
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/r02_renderer_variants.jsonl
for sfx in _rb _rt "" _ru3; do
  [ -f mistral-water_b200/lib/libmistral_ocean$sfx.so ] || continue
  MW_LIB_SUFFIX=$sfx timeout 300 python tools/bench_extra.py --only renderer 2>>$OUT/r02_renderer_variants.err | sed "s/^{/{\"lib\": \"$sfx\", /" >> $OUT/r02_renderer_variants.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_renderer_variants.jsonl'):
    d=json.loads(l); print(repr(d['lib']).ljust(7), d['texture_resolution'], d['tiles'], d['us_per_frame'], d['frac_of_measured_hbm'])
PY
