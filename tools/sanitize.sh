#!/bin/bash
# Runs on the GPU box: compute-sanitizer memcheck + racecheck over one small frame of every path (developer check).
set -u
cat > /tmp/san.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import mistral_water_b200 as mw
for N in (32, 64, 256, 1024):
    with mw.Ocean(N, seed=3, tiles=2) as o:
        o.init_spectrum()
        o.generate(0.7, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
        o.generate(0.9, names=("height", "disp", "normal"))
with mw.Ocean(2048, seed=3) as o:
    o.init_spectrum(); o.generate(0.7)
for res in (4, 32, 128):
    with mw.Renderer(res, 434.48, 0.46, 0.41, (14.45, 12.0), 1.5, tiles=2) as r:
        r.render_initial(); r.generate_texture(0.016, names=("displacement", "height", "normal", "white", "white_rgba", "jacobian"))
g = mw.pond_wave_table_32()
pos = np.random.default_rng(0).uniform(-50, 50, (1001, 3)).astype(np.float32)
g.displace(pos, 1.0); mw.wave_displace(pos, 1.0, 10, 2.5, 1.3, 0.4); mw.generate_mesh(33, 1.0)
x = (np.random.default_rng(1).standard_normal((2, 64, 64)) + 0j).astype(np.complex64); mw.fft2d(x)
print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/san.py 2>&1 | grep -v "^=========     Saved|Host Frame|^=========         in|^=========                in" | head -60
  echo "exit $?"
done
