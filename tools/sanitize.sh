#!/bin/bash
# Runs on the GPU box: compute-sanitizer memcheck + racecheck + synccheck over one small frame of every path (developer check;
# SURVEY.md section 5.2).  Usage: bash tools/sanitize.sh > gpurun_out/r02_sanitizer.txt
set -u
cat > /tmp/san.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import mistral_water_b200 as mw
ONLY_TILES = os.environ.get("MW_SAN_ONLY", "") == "tiles"    # only the tile-set part (the 2-GPU re-check of the push engines)
for N in (() if ONLY_TILES else (32, 64, 256, 1024)):
    with mw.Ocean(N, seed=3, tiles=2) as o:
        o.init_spectrum()
        o.generate(0.7, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
        o.generate(0.9, names=("height", "disp", "normal"))
if not ONLY_TILES:
    with mw.Ocean(2048, seed=3) as o:
        o.init_spectrum(); o.generate(0.7)
for res in (() if ONLY_TILES else (4, 32, 128)):
    with mw.Renderer(res, 434.48, 0.46, 0.41, (14.45, 12.0), 1.5, tiles=2) as r:
        r.render_initial(); r.generate_texture(0.016, names=("displacement", "height", "normal", "white", "white_rgba", "jacobian"))
for N, L in (() if ONLY_TILES else ((12, 12.39), (13, 13.0), (16, 16.0))):     # direct-sum path (the FFT Mesh scene's own grid)
    with mw.Ocean(N, length=L, seed=3, tiles=2) as o:
        o.init_spectrum()
        o.generate(0.7, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
for N in (() if ONLY_TILES else (64, 1024)):            # graph replay: third call with the same (scratch) outputs
    with mw.Ocean(N, seed=4) as o:
        o.init_spectrum()
        for k in range(4):
            o.generate(0.1 * k)
from mistral_water_b200.tiles import TileSet
with TileSet(64, 1, rank=None, tiles_per_rank=2, gather="peer", asynchronous=False) as ts:
    ts.init_spectrum()
    for k in range(3):
        ts.generate_allgather(0.1 * k)
import torch
if torch.cuda.device_count() >= 2:                     # the peer arm's three push engines, one process driving two GPUs
    for push in ("tma", "sm", "ce"):
        with TileSet(64, 2, rank=None, devices=[0, 1], tiles_per_rank=2, gather="peer", asynchronous=True, push=push) as ts:
            ts.init_spectrum()
            for k in range(4):
                ts.generate_allgather(0.1 * k)
            ts.sync()
    print("two-GPU tile sets done")
g = mw.pond_wave_table_32()
pos = np.random.default_rng(0).uniform(-50, 50, (1001, 3)).astype(np.float32)
g.displace(pos, 1.0); g.displace(pos, 1.0, normals=np.empty_like(pos), normal_mode="analytic")
g.displace(pos, 1.0, normals=np.empty_like(pos), normal_mode="discarded", smoothing=0.4); mw.wave_displace(pos, 1.0, 10, 2.5, 1.3, 0.4); mw.generate_mesh(33, 1.0)
x = (np.random.default_rng(1).standard_normal((2, 64, 64)) + 0j).astype(np.complex64); mw.fft2d(x)
print("sanitizer workload done")
PY
for tool in ${MW_SAN_TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/san.py 2>&1 | grep -v "^=========     Saved|Host Frame|^=========         in|^=========                in" | head -60
  echo "exit $?"
done
