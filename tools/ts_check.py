import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import mistral_water_b200 as mw
N = 1024
with mw.Ocean(N, seed=7, tiles=3) as o:
    o.init_spectrum()
    a = o.generate(0.8, names=("height", "disp", "normal", "whitecap"))              # OUTS = 7 path (TMA stores in the _ts build)
    b = o.generate(0.8, names=("height", "disp", "normal", "whitecap", "jacobian"))  # run-time output set: per-thread stores
for k in a:
    print(k, np.array_equal(a[k], b[k]), float(np.abs(a[k] - b[k]).max()))
