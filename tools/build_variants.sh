#!/bin/bash
# Runs here (no GPU): experiment builds next to the product library (they travel to the GPU box with the snapshot).
#   _occ : occupancy of the resolutions the bench does not run
set -eu
cd "$(dirname "$0")/.."
MW_LIB_SUFFIX=_occ MW_NVCC_DEFS="-DMW_COLS_MAXREG_512=112 -DMW_COLS_MAXREG_256=96 -DMW_ROWS_MINB_256=3 -DMW_ROWS_MINB_512=3 -DMW_ROWS_MINB_2048=2" \
    python mistral-water_b200/build.py
