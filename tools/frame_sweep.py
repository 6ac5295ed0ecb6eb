"""16 x 1024^2 frame + single-frame time + per-kernel times (MW_PROFILE, serialised) of the library build MW_LIB_SUFFIX selects,
under the environment's MW_* switches.  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mistral_water_b200 as mw

st = torch.cuda.Stream()
COMPS = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}
N = int(os.environ.get("MW_SWEEP_N", "1024")); TILES = int(os.environ.get("MW_SWEEP_TILES", "16"))


def frame_us(N, tiles, K, reps=3, profile=False):
    o = mw.Ocean(N, seed=1000, tiles=tiles, device_ptrs=True, profile=profile)
    o.set_stream(st.cuda_stream); o.init_spectrum()
    n2 = N * N * tiles
    bufs = {k: torch.empty(n2 * c, device="cuda") for k, c in COMPS.items()}
    best, kt = 1e30, None
    with torch.cuda.stream(st):
        for i in range(10): o.generate(0.016 * i, bufs)
        torch.cuda.synchronize()
        if profile:
            o.kernel_times(reset=True)
            for i in range(K): o.generate(0.016 * i, bufs)
            ms, n = o.kernel_times()
            kt = [round(ms[i] / K * 1e3, 1) for i in range(2)]
        else:
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for i in range(K): o.generate(0.016 * i, bufs)
                e1.record(st); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K * 1e3)
    o.close()
    return kt if profile else round(best, 2)


res = {"lib": os.environ.get("MW_LIB_SUFFIX", ""), "env": {k: v for k, v in os.environ.items() if k.startswith("MW_") and k not in ("MW_LIB_SUFFIX",)}}
res[f"{TILES}x{N}"] = frame_us(N, TILES, 40)
res[f"1x{N}"] = frame_us(N, 1, 200)
res["rows_cols_us_per_frame_serialised"] = frame_us(N, TILES, 20, profile=True)
print(json.dumps(res), flush=True)
