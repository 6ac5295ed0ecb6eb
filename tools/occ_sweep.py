"""Frame times at the resolutions other than the bench's (occupancy experiments; MW_LIB_SUFFIX picks the library build)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mistral_water_b200 as mw

st = torch.cuda.Stream()
ALL = ("height", "disp", "normal", "whitecap")
COMPS = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}


def frame_us(N, tiles, names=ALL, K=50, reps=3):
    o = mw.Ocean(N, seed=1000, tiles=tiles, device_ptrs=True)
    o.set_stream(st.cuda_stream); o.init_spectrum()
    n2 = N * N * tiles
    bufs = {k: torch.empty(n2 * COMPS[k], device="cuda") for k in names}
    best = 1e30
    with torch.cuda.stream(st):
        for i in range(10): o.generate(0.016 * i, bufs)
        torch.cuda.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(K): o.generate(0.016 * i, bufs)
            e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K * 1e3)
    o.close()
    return round(best, 2)


res = {"lib": os.environ.get("MW_LIB_SUFFIX", ""), "env": {k: v for k, v in os.environ.items() if k.startswith("MW_") and k != "MW_LIB_SUFFIX"}}
res["256x256_hdn"] = frame_us(256, 256, ("height", "disp", "normal"), K=20)
res["256x256_all"] = frame_us(256, 256, K=20)
res["64x512"] = frame_us(512, 64, K=20)
res["1x512"] = frame_us(512, 1, K=200)
res["4x2048"] = frame_us(2048, 4, K=20)
res["1x2048"] = frame_us(2048, 1, K=50)
res["1024x128"] = frame_us(128, 1024, K=20)
print(json.dumps(res), flush=True)
