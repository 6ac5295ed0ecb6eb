"""Secondary measurements (not the driver's bench line): the other BASELINE configs on one GPU.
Prints one JSON object per config: 256^2 (config 2, single + batched), 1024^2 single tile latency (config 3 as the
C# host would call it), 2048^2 tile (config 5's per-rank work), Gerstner 32 waves x 1M vertices (config 4)."""
import json, os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mistral_water_b200 as mw

ap = argparse.ArgumentParser(); ap.add_argument("--only", default=""); args = ap.parse_args()
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
st = torch.cuda.Stream()


def ocean(N, tiles, names, K=100, label=""):
    o = mw.Ocean(N, seed=1000, tiles=tiles, device_ptrs=True)
    o.set_stream(st.cuda_stream); o.init_spectrum()
    n2 = N * N * tiles
    comps = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}
    bufs = {k: torch.empty(n2 * comps[k], device="cuda") for k in names}
    bpp = 16 + 4 * sum(comps[k] for k in names)
    with torch.cuda.stream(st):
        for i in range(10): o.generate(0.016 * i, bufs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(K): o.generate(0.016 * i, bufs)
        e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    o.close()
    print(json.dumps({"config": label, "resolution": N, "tiles": tiles, "outputs": list(names), "us_per_frame": round(ms * 1e3, 2),
                      "grid_points_per_s": n2 / ms * 1e3, "algorithmic_bytes_per_point": bpp,
                      "achieved_gbs": round(n2 * bpp / ms / 1e6, 1), "frac_of_measured_hbm": round(n2 * bpp / ms / 1e6 / PEAK, 4),
                      "l2_note": "working set fits L2: a latency number, not an HBM number" if n2 * (bpp + 28) < 100e6 else "exceeds L2"}), flush=True)


def gerstner(K=50):
    N = 1024
    g = mw.pond_wave_table_32(device_ptrs=True)
    ax = (torch.arange(N, device="cuda", dtype=torch.float32) - N // 2 + 0.5)
    pos = torch.zeros(N * N, 3, device="cuda"); pos[:, 0] = ax.repeat_interleave(N); pos[:, 2] = ax.repeat(N)
    out = torch.empty_like(pos)
    flush = torch.empty(64 << 20, device="cuda")  # 256 MB > L2
    with torch.cuda.stream(st):
        for i in range(5): g.displace(pos, 1.7, out=out, stream=st.cuda_stream)
        torch.cuda.synchronize()
        tot = 0.0
        for i in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.displace(pos, 1.7 + 0.016 * i, out=out, stream=st.cuda_stream); e1.record(st)
            torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms = tot / K
    print(json.dumps({"config": "4: Gerstner 32 waves x 1048576 vertices", "us": round(ms * 1e3, 2), "vertices_per_s": N * N / ms * 1e3,
                      "algorithmic_bytes_per_vertex": 24, "achieved_gbs": round(N * N * 24 / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(N * N * 24 / ms / 1e6 / PEAK, 4), "l2": "flushed between iterations (256 MB memset)",
                      "bound": "MUFU/FP32 issue (64 transcendentals per vertex), not HBM"}), flush=True)


def renderer(res, tiles, K=50, label=""):
    """OceanRenderer path: GenerateTexture -> displacement, height, normal (RGBAFloat) + white (R channel)."""
    R = 8 * res
    r = mw.Renderer(res, 434.48, 0.46, 0.41, (14.45, 12.0), 1.5, seed1=3.7, seed2=8.1, tiles=tiles, device_ptrs=True)
    r.render_initial()
    n2 = R * R * tiles
    bufs = {"displacement": torch.empty(n2 * 4, device="cuda"), "height": torch.empty(n2 * 4, device="cuda"),
            "normal": torch.empty(n2 * 4, device="cuda"), "white": torch.empty(n2, device="cuda")}
    bpp = 16 + 4 + 4 + 16 + 16 + 16 + 4  # read initial + phase, write phase + three RGBA maps + white.r
    for i in range(5): r.generate_texture(0.016, bufs)
    r.sync(); torch.cuda.synchronize()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the handle owns its stream: bracket with host sync + device events on the default stream would not see it,
    # so time K frames between two mw_renderer_sync calls (K large enough that launch latency is hidden)
    t0 = time.perf_counter()
    for i in range(K): r.generate_texture(0.016, bufs)
    r.sync()
    ms = (time.perf_counter() - t0) * 1e3 / K
    r.close()
    print(json.dumps({"config": label, "texture_resolution": R, "mesh_resolution": res, "tiles": tiles,
                      "maps": "displacement,height,normal (RGBAFloat) + white.r", "us_per_frame": round(ms * 1e3, 2),
                      "texels_per_s": n2 / ms * 1e3, "algorithmic_bytes_per_texel": bpp, "achieved_gbs": round(n2 * bpp / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(n2 * bpp / ms / 1e6 / PEAK, 4), "timing": "wall clock between two stream syncs over K frames",
                      "reference": "44 + 44 + 5 full-texture blits per frame at 16-48 B/texel each (OceanRenderer.cs:216-316)"}), flush=True)


if args.only == "renderer16":
    renderer(128, 16, K=4, label="Ocean Demo scene x 16 oceans per call (profiling run)")
if args.only in ("", "renderer"):
    renderer(128, 1, label="Ocean Demo scene: resolution 128 -> 1024^2 maps, one ocean per call")
    renderer(128, 16, K=20, label="Ocean Demo scene x 16 oceans per call")
    renderer(256, 1, label="OceanRenderer default: resolution 256 -> 2048^2 maps")
    renderer(256, 4, K=20, label="OceanRenderer default x 4 oceans per call")
if args.only in ("", "ocean"):
    ocean(64, 1, ("height", "disp", "normal", "whitecap"), label="1: 64x64 single tile (plumbing; launch-latency bound)")
    ocean(256, 1, ("height", "disp", "normal"), label="2: 256x256 height+disp+normal, single tile (launch-latency bound)")
    ocean(256, 256, ("height", "disp", "normal"), K=30, label="2: 256x256 height+disp+normal, 256 tiles per call")
    ocean(1024, 1, ("height", "disp", "normal", "whitecap"), label="3: 1024x1024 + whitecap, single tile per call (what one FFTMesh does)")
    ocean(2048, 1, ("height", "disp", "normal", "whitecap"), K=50, label="5: one 2048x2048 tile (per-rank work of config 5)")
    ocean(2048, 4, ("height", "disp", "normal", "whitecap"), K=30, label="5: four 2048x2048 tiles per call")
if args.only in ("", "gerstner"):
    gerstner()
