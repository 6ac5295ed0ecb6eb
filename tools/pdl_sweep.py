"""Frame time of the FFTMesh path under the environment's launch options (MW_PDL, MW_ROWS_MINB, ...), one JSON line:
16 x 1024^2 per call (the bench workload, default tile-group scheduling) and the single-tile latencies."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mistral_water_b200 as mw

st = torch.cuda.Stream()
ALL = ("height", "disp", "normal", "whitecap")
COMPS = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}


def frame_us(N, tiles, names=ALL, K=200, reps=3):
    o = mw.Ocean(N, seed=1000, tiles=tiles, device_ptrs=True)
    o.set_stream(st.cuda_stream); o.init_spectrum()
    n2 = N * N * tiles
    bufs = {k: torch.empty(n2 * COMPS[k], device="cuda") for k in names}
    best = 1e30
    with torch.cuda.stream(st):
        for i in range(20): o.generate(0.016 * i, bufs)
        torch.cuda.synchronize()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(K): o.generate(0.016 * i, bufs)
            e1.record(st); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K * 1e3)
    o.close()
    return round(best, 2)


res = {"env": {k: v for k, v in os.environ.items() if k.startswith("MW_")}}
res["16x1024"] = frame_us(1024, 16, K=50)
res["1x1024"] = frame_us(1024, 1)
res["1x2048"] = frame_us(2048, 1, K=50)
res["1x256_hdn"] = frame_us(256, 1, ("height", "disp", "normal"))
res["1x64"] = frame_us(64, 1)
res["256x256_hdn"] = frame_us(256, 256, ("height", "disp", "normal"), K=20)
print(json.dumps(res), flush=True)
