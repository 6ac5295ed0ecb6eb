"""One 16 x 1024^2 frame of the bench workload in the TIMED scheduling (one tile per group, two streams), for
`ncu --replay-mode application --cache-control none`: DRAM bytes of every frame kernel of ONE whole mw_ocean_generate,
summed by tools/summarize_traffic.py.  Warm-up: W frames = W * (1 + 2 * tiles) launches to skip with -s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mistral_water_b200 as mw

N = int(os.environ.get("MW_TR_N", "1024")); tiles = int(os.environ.get("MW_TR_TILES", "16")); W = int(os.environ.get("MW_TR_WARM", "3"))
st = torch.cuda.Stream()
o = mw.Ocean(N, seed=1000, tiles=tiles, device_ptrs=True)
o.set_stream(st.cuda_stream); o.init_spectrum()
n2 = N * N * tiles
bufs = {"height": torch.empty(n2, device="cuda"), "disp": torch.empty(n2 * 2, device="cuda"),
        "normal": torch.empty(n2 * 3, device="cuda"), "whitecap": torch.empty(n2, device="cuda")}
with torch.cuda.stream(st):
    for i in range(W + 1):
        o.generate(0.016 * i, bufs)
    torch.cuda.synchronize()
o.close()
