#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

N = 1 (BASELINE.json configs[2], the one the metric's "disp+Jacobian" is quoted on):
    one step = one frame of a batch of 16 independent 1024 x 1024 Tessendorf grids, each
    h0 -> h(k,t) -> 2-D IFFT -> height + hds + normal + Jacobian whitecap (44 algorithmic B/point),
    outputs left in HBM.  16 tiles make the per-step input (268 MB) and output (470 MB) larger
    than the 126 MB L2, so no flush is needed between timed steps.
N > 1 (torchrun, one rank per GPU; BASELINE.json configs[4]):
    every rank generates ONE 2048 x 2048 tile (seed 1000 + rank, wind rotated 45 deg * rank) and the step
    ends with the path's one collective, the in-place all-gather of the final float buffers
    (117.4 MB per rank), through the mw_tiles_* C ABI: peer-memory pushes (default arm) and ncclAllGather
    (second arm, reported beside it).  Every rank issues exactly the same sequence of collective calls.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # the CUDA engine
    python bench.py --impl reference [...]                       # the reference's CPU algorithm (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-points/sec (spectrum→IFFT→disp+Jacobian) at N×N; achieved HBM GB/s"
UNIT = "grid-points/s"
ALG_BYTES_PIPELINE = 44  # read h0+h0conj 16, write height 4 + hds 8 + normal 12 + whitecap 4
ALG_BYTES_KERNEL = {"spectrum_rows": 16 + 24, "cols": 24 + 28}  # a kernel's own bytes: + the 24 B/pt intermediate
NVLINK_NOMINAL_GBS, NVLINK_PEER_COPY_GBS = 900.0, 770.0   # per direction per GPU: nominal / measured peer copy (B200_PROFILING.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks during the timed region
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples, self.reasons, self.power = [], set(), []
        self.period, self._stop, self._thr, self.max_mhz = period_s, threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": round(max(self.power), 1) if self.power else None}


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_rate(N, seed, vertices, threads, t=1.7):
    """Literal O(N^2)-per-vertex FFTMesh.Displacement on `vertices` vertices of the N x N grid."""
    import numpy as np
    from oracle import cref

    p = cref.params(N)
    v, h0, hc = cref.generate_mesh(p, seed=seed)
    start = (N * N) // 2 - vertices // 2
    t0 = time.perf_counter()
    cref.evaluate_vertices(p, v, h0, hc, t, start, start + vertices, threads=threads)
    dt = time.perf_counter() - t0
    return vertices / dt, dt


def cpu_fft_form_rate(N, seed, t=1.7):
    """The same frame on the CPU in TRANSFORM form (oracle/ref_fft64.py: numpy fp64, five ifft2 + the extraction of
    FFTMesh.cs:243-276), one thread: separates what the algorithm buys (O(N^2 log N) instead of the reference's
    O(N^4)) from what the hardware buys.  Not what the reference does -- labelled as such in the JSON line."""
    from oracle import cref, ref_fft64

    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=seed)
    ref_fft64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, 0.0)  # warm-up (FFT plans, page faults)
    t0 = time.perf_counter()
    ref_fft64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, t)
    dt = time.perf_counter() - t0
    return N * N / dt, dt


def defaults_for_world(args, world):
    """N = 1: configs[2] batched (16 x 1024^2).  N > 1: configs[4] (one 2048^2 tile per GPU)."""
    if args.resolution is None:
        args.resolution = 1024 if world == 1 else 2048
    if args.tiles is None:
        args.tiles = 16 if world == 1 else 1
    return args


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)   # the workload is the one the engine arm runs at this N
    if rank != 0:
        return
    defaults_for_world(args, world)
    from oracle import cref
    cref.build()
    N = args.resolution
    # all host threads, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_vertex_s = 1.0e-7 * N * N  # ~100 ns per inner term
    verts = max(threads, int(round(args.ref_step_seconds * threads / per_vertex_s)))
    for _ in range(args.warmup):
        cpu_reference_rate(N, 1000, max(threads, verts // 4), threads)
    total_v, total_t = 0, 0.0
    for k in range(args.steps):
        r, dt = cpu_reference_rate(N, 1000, verts, threads, t=0.016 * k)
        total_v += verts
        total_t += dt
    value = total_v / total_t
    sample = (f"{verts} of {N * N} vertices per step through the literal FFTMesh.Displacement loop "
              f"(N^2 = {N * N} wave vectors each), {threads} OpenMP threads over vertices; Unity itself runs "
              f"this on one thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    N, T = args.resolution, args.tiles
    if world == 1:
        name = (f"{T} x ({N}x{N} Tessendorf grid, height+hds+normal+Jacobian whitecap) per step"
                + (" = BASELINE configs[2] batched" if N == 1024 else ""))
        coll = "none"
    else:
        name = (f"{world} x {T} independent {N}x{N} ocean tiles, {T} per GPU (seed 1000+tile, wind rotated 45 deg per rank), "
                f"all-gather of the final float buffers ({T * N * N * 28 / 1e6:.1f} MB per rank)"
                + (" = BASELINE configs[4]" if (N, T) == (2048, 1) else ""))
        coll = ("one in-place all-gather of the final float buffers per step through the tile-set handle (mw_tiles_*), "
                "MW_GATHER_AUTO: every rank pushes its slot into the peers' buffers with a TMA bulk-copy kernel (CUDA IPC mappings, fenced "
                "by stream memory operations); the ncclAllGather arm is measured beside it and both are reported under multi_gpu")
    return {
        "workload": name,
        "resolution": N, "tiles_per_gpu": T, "points_per_step_per_gpu": T * N * N,
        "outputs": "height,hds,normal,whitecap (28 B/pt)", "algorithmic_bytes_per_point": ALG_BYTES_PIPELINE,
        "l2": f"inputs {T * N * N * 16 / 1e6:.0f} MB + outputs {T * N * N * 28 / 1e6:.0f} MB per step > 126 MB L2; no flush",
        "collective": coll,
        "parallelism": f"tiles{world}",
    }


# ----------------------------------------------------------------------------- the other BASELINE configs (rank 0, N = 1)
def extra_configs(mw, torch, stream, peak):
    """Driver-visible lines for BASELINE configs 1, 2, 4 (Gerstner), the single 1024^2 frame and the 2048^2 tile:
    CUDA events on the launching stream, 10 warm-up frames.  Single small frames fit L2: latency numbers, labelled."""
    out = {}
    comps = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}

    def ocean(key, N, tiles, names, K):
        o = mw.Ocean(N, seed=1234 if N <= 256 else 1000, tiles=tiles, device_ptrs=True)
        o.set_stream(stream.cuda_stream)
        o.init_spectrum()
        n2 = N * N * tiles
        bufs = {k: torch.empty(n2 * comps[k], device="cuda") for k in names}
        bpp = 16 + 4 * sum(comps[k] for k in names)
        with torch.cuda.stream(stream):
            for i in range(10):
                o.generate(0.016 * i, bufs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(K):
                o.generate(0.016 * i, bufs)
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        o.close()
        fits = n2 * (bpp + 24) < 100e6
        out[key] = {"resolution": N, "tiles_per_call": tiles, "outputs": list(names), "us_per_frame": round(ms * 1e3, 2),
                    "value": n2 / ms * 1e3, "unit": UNIT, "algorithmic_bytes_per_point": bpp,
                    "achieved_gbs": round(n2 * bpp / ms / 1e6, 1), "frac_of_hbm_peak": round(n2 * bpp / ms / 1e6 / peak, 4),
                    "l2": "working set fits the 126 MB L2: a latency figure, not an HBM figure" if fits else "working set exceeds L2"}

    all4, hdn = ("height", "disp", "normal", "whitecap"), ("height", "disp", "normal")
    ocean("config1_64x64_single_frame", 64, 1, all4, 200)
    ocean("config2_256x256_single_frame", 256, 1, hdn, 200)
    ocean("config2_256x256_x256_per_call", 256, 256, hdn, 20)
    ocean("config3_1024x1024_single_frame", 1024, 1, all4, 200)
    ocean("config5_one_2048x2048_tile", 2048, 1, all4, 50)
    ocean("config5_four_2048x2048_tiles", 2048, 4, all4, 20)
    # the reference's own FFT Mesh demo scene (12 x 12, length 12.39: direct-sum kernels) -- what its CPU loop is actually run on
    o = mw.Ocean(12, length=12.39, seed=1234, device_ptrs=True)
    o.set_stream(stream.cuda_stream)
    o.init_spectrum()
    bufs = {k: torch.empty(144 * comps[k], device="cuda") for k in all4}
    with torch.cuda.stream(stream):
        for i in range(10):
            o.generate(0.016 * i, bufs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(200):
            o.generate(0.016 * i, bufs)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    o.close()
    out["fft_mesh_demo_scene_12x12_direct_sum"] = {
        "resolution": 12, "length": 12.39, "us_per_frame": round(ms * 1e3, 2), "value": 144 / ms * 1e3, "unit": UNIT,
        "note": "Demo/FFT Mesh.unity:145-152 as shipped: not periodic, so the O(N^4) sum itself runs on the GPU (three launches; a latency figure)"}
    # config 4: Gerstner 32 waves x 1M vertices, L2 flushed between iterations (the 24 MB working set would sit in L2)
    N = 1024
    g = mw.pond_wave_table_32(device_ptrs=True)
    ax = torch.arange(N, device="cuda", dtype=torch.float32) - N // 2 + 0.5
    pos = torch.zeros(N * N, 3, device="cuda")
    pos[:, 0] = ax.repeat_interleave(N)
    pos[:, 2] = ax.repeat(N)
    res = torch.empty_like(pos)
    flush = torch.empty(64 << 20, device="cuda")
    with torch.cuda.stream(stream):
        for i in range(5):
            g.displace(pos, 1.7, out=res, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        tot, K = 0.0, 30
        for i in range(K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.displace(pos, 1.7 + 0.016 * i, out=res, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    ms = tot / K
    out["config4_gerstner_32_waves_1M_vertices"] = {
        "us": round(ms * 1e3, 2), "value": N * N / ms * 1e3, "unit": "vertices/s", "algorithmic_bytes_per_vertex": 24,
        "achieved_gbs": round(N * N * 24 / ms / 1e6, 1), "frac_of_hbm_peak": round(N * N * 24 / ms / 1e6 / peak, 4),
        "l2": "flushed between iterations (256 MB memset)",
        "bound": "MUFU/FP32 issue (64 transcendentals per vertex at 24 B/vertex), not HBM: profiles/ has the pipe utilisation"}
    return out


def bind_to_gpu_numa(index):
    """Best effort: run this rank on the CPUs NVML names as local to its GPU, so that its pinned host buffers are first
    touched on that NUMA node (round 1: per-rank e2e degraded 9.8 -> 49 ms at 8 ranks with every rank on CPUs 0-31)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = ideal & allowed
        if not use:
            return {"bound": False, "why": "the GPU's local CPUs are outside this process' allowed set", "allowed": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"bound": True, "cpus": len(use), "first_cpu": min(use)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": repr(e)}


# ----------------------------------------------------------------------------- the CUDA engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist

    import mistral_water_b200 as mw
    from mistral_water_b200.tiles import FIELDS, ShardedTiles

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else None     # before any pinned allocation: first touch decides the node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    defaults_for_world(args, world)
    N, T, K, W = args.resolution, args.tiles, args.steps, max(args.warmup, 3)
    pts_rank = T * N * N
    peak, peak_src = peaks()
    slot_bytes = pts_rank * 28

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(st, fn, n, finish=True):
        """n calls of fn(k) between two CUDA events on the user stream; every rank makes the same calls."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st.stream)
        for k in range(n):
            fn(k)
        if finish:
            st.finish()
        e1.record(st.stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def open_tiles(gather):
        """The tile set of one gather arm, or None when ANY rank could not set it up (e.g. a box whose containers cannot share CUDA
        IPC handles): the verdict is all-reduced, so every rank takes the same branch afterwards.  Returns (tiles, error text)."""
        st, err = None, ""
        try:
            st = ShardedTiles(N, rank, world, tiles_per_rank=T, base_seed=1000, device=dev, gather=gather)
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
        if world > 1:
            ok = torch.tensor([0 if st is None else 1], device=dev, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if st is not None:
                    st.tileset.close()      # no collective teardown: the peers never connected
                    st.tileset = None
                return None, err or "another rank failed to set up this arm"
        elif st is None:
            raise RuntimeError(err)
        return st, ""

    def measure_arm(gather, st=None):
        """One gather arm at world > 1 (or the plain engine at world == 1): the timed step, then compute alone and the
        gather alone.  Rank-uniform control flow throughout: K, W and the leg order are the same on every rank."""
        if st is None:
            st, err = open_tiles(gather)
            if st is None:
                return None, {"impl": gather, "unavailable": err}
        step = (lambda k: st.generate_pipelined(0.016 * k)) if world > 1 else (lambda k: st.generate_local(0.016 * k))
        res = {"impl": st.gather_impl}
        with torch.cuda.stream(st.stream):
            for k in range(W):
                step(k)
            st.finish()
            barrier()
            launches0 = mw.native.launch_count()
            with ClockSampler(local) as clk:
                res["ms"] = timed(st, lambda k: step(W + k), K)
                res["launches"] = mw.native.launch_count() - launches0
                # a short timed region gives the sampler too few looks: keep the SAME KERNELS running untimed -- compute
                # only (generate_local has no collective), so a rank-local number of extra iterations cannot desynchronise
                # the ranks
                extra = 0
                while len(clk.samples) < 8 and extra < 200:
                    for k in range(8):
                        st.generate_local(0.016 * k)
                    st.finish()
                    torch.cuda.synchronize()
                    extra += 1
            res["clocks"] = clk.summary()
            res["clocks"]["sampled"] = "timed region" + (f" + {extra * 8} untimed compute-only steps of the same kernels" if extra else "")
            res["compute_ms"] = res["ms"]
            res["gather_ms"] = 0.0
            if world > 1:
                res["compute_ms"] = timed(st, lambda k: st.generate_local(0.016 * k), K)
                st.generate_local(0.0)
                st.finish()
                res["gather_ms"] = timed(st, lambda k: st.all_gather(), K)
        st.sync()
        return st, res

    st, main = measure_arm("auto")     # MW_GATHER_AUTO: the peer pushes (include/mistral_ocean.h)
    auto_fallback = None
    if st is None:
        # the default arm could not be set up on this box: the step is measured with ncclAllGather instead, and the line says so
        auto_fallback = main["unavailable"]
        st, main = measure_arm("nccl")
        if st is None:
            raise RuntimeError(f"neither gather arm could be set up: {auto_fallback}; {main['unavailable']}")
    ms, compute_ms, clocks, launches = main["ms"], main["compute_ms"], main["clocks"], main["launches"]
    value = world * pts_rank * K / (ms * 1e-3)
    stream = st.stream

    # ---- roofline: the SURVEY 8(d) figure -- 44 algorithmic B/pt x points per step / step time of the timed region ----
    roof = None
    if rank == 0:
        frame_ms = compute_ms / K
        ach = ALG_BYTES_PIPELINE * pts_rank / (frame_ms * 1e-3) / 1e9
        roof = {
            "bound": "hbm", "kernel": "k_spectrum_rows + k_cols_seam (one frame = one mw_ocean_generate)",
            "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_point": ALG_BYTES_PIPELINE, "points_per_step": pts_rank,
            "algorithmic_bytes_per_step": ALG_BYTES_PIPELINE * pts_rank,
            "timing": "CUDA events on the launching stream around the K timed steps (compute only), max over ranks",
        }
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                t = json.load(open(tr)).get(f"{N}x{T}")
                if t:
                    roof["traffic"] = t["dram_bytes_per_step"]
                    roof["traffic_over_algorithmic"] = round(t["dram_bytes_per_step"] / (ALG_BYTES_PIPELINE * pts_rank), 3)
                    roof["traffic_source"] = t["source"]
            except Exception:  # noqa: BLE001
                pass
        # per-kernel durations, live, with CUDA events on the launching stream (MW_PROFILE handle: single stream, so the
        # launches do not overlap -- shares of the step, not timings of the timed region)
        prof = mw.Ocean(N, seed=1000, tiles=T, device=local, device_ptrs=True, profile=True)
        prof.set_stream(stream.cuda_stream)
        prof.init_spectrum()
        views = {k: torch.empty(pts_rank * c, device=dev) for k, c in FIELDS}
        with torch.cuda.stream(stream):
            for k in range(3):
                prof.generate(0.016 * k, views)
            prof.sync()
            prof.kernel_times(reset=True)
            for k in range(K):
                prof.generate(0.016 * k, views)
            kms, kn = prof.kernel_times()
        prof.close()
        del views
        names = ["spectrum_rows", "cols"]   # pass 2 is k_cols_seam from N = 512 up, k_cols_extract below
        per = {names[i]: kms[i] / max(kn[i], 1) for i in range(2)}
        per_step = {names[i]: kms[i] / K for i in range(2)}
        lps = {names[i]: kn[i] / K for i in range(2)}
        roof["kernels"] = {
            "k_" + nm: {"avg_launch_ms": round(per[nm], 4), "launches_per_step": lps[nm],
                        "share_of_step": round(per_step[nm] / sum(per_step.values()), 3),
                        "own_bytes_per_point": ALG_BYTES_KERNEL[nm],
                        "own_gbs": round(ALG_BYTES_KERNEL[nm] * pts_rank / lps[nm] / (per[nm] * 1e-3) / 1e9, 1)}
            for nm in names}
        roof["kernels"]["note"] = ("serialised on one stream (MW_PROFILE); own_bytes include the 24 B/pt intermediate the kernel "
                                   "itself moves (L2-resident in the timed scheduling) -- sub-figures, not the roofline fraction")

    # ---- the other gather arm beside the default one (world > 1): same legs, same call counts on every rank ----
    other = None
    if world > 1:
        st.close()
        barrier()
        if auto_fallback is None:
            st2, other = measure_arm("nccl" if main["impl"] == "peer" else "peer")
            if st2 is not None:
                st2.close()
        else:
            other = {"impl": "peer", "unavailable": auto_fallback}
        barrier()
        st, _ = None, None

    # ---- end to end through the C ABI with HOST buffers ----
    Ke = max(1, min(K, args.e2e_steps))
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float32).pin_memory()  # noqa: E731
    if world == 1:
        # what the C# host calls: mw_ocean_set_h0 + mw_ocean_generate on pinned host arrays.  MW_HOST_ASYNC: calls enqueue and
        # return; results leave on the handle's copy stream, so the upload of step k + 1 overlaps the download of step k
        host = mw.Ocean(N, seed=1000, tiles=T, device=local, host_async=True)
        h0, h0c = pin(pts_rank, 2), pin(pts_rank, 2)
        host.init_spectrum()
        host.get_h0_into(h0, h0c)
        outs = {name: pin(pts_rank, c) for name, c in FIELDS}
        for k in range(2):
            host.set_h0(h0, h0c)
            host.generate(0.016 * k, outs)
        host.sync()
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            host.set_h0(h0, h0c)                 # verttilde / vertConj from host memory, every call
            host.generate(0.016 * k, outs)       # results land in host arrays (complete at sync)
        host.sync()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        checksum = float(outs["height"][: N * N].double().abs().sum())
        host.close()
        e2e = {"value": world * pts_rank * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pts_rank * 16,
               "d2h_bytes_per_step": pts_rank * 28, "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
               "api": "mw_ocean_set_h0 + mw_ocean_generate with pinned host buffers on an MW_HOST_ASYNC handle (upload of step "
                      "k+1 overlaps download of step k), mw_ocean_sync at the end", "height_abs_sum_tile0": checksum}
    else:
        # the tile set: every step uploads this rank's h0 from pinned host memory, generates, ALL-GATHERS, and downloads this
        # rank's slot of the gathered buffer once the gather has completed (the hosts fetch each tile once, over N PCIe links)
        st3 = ShardedTiles(N, rank, world, tiles_per_rank=T, base_seed=1000, device=dev, gather=main["impl"])   # the arm `value` was measured with
        ts = st3.tileset
        h0, h0c = pin(pts_rank, 2), pin(pts_rank, 2)
        d_h0 = [torch.empty(pts_rank, 2, device=dev) for _ in range(2)]      # upload staging, double-buffered
        d_h0c = [torch.empty(pts_rank, 2, device=dev) for _ in range(2)]
        lib = mw.native.load()
        mw.native.check(lib.mw_ocean_get_h0(ts.ocean_handle(0), d_h0[0].data_ptr(), d_h0c[0].data_ptr()))
        ts.sync()
        h0.copy_(d_h0[0])
        h0c.copy_(d_h0c[0])
        out_host = pin(pts_rank * 7)
        d2h = torch.cuda.Stream(device=dev)
        ev_gathered, ev_fetched = [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]

        def e2e_step(k):
            # Uploads go on the user stream, after the previous frame's gather (measured on 2 x B200: 3.06 ms per step; running
            # them ahead on a separate stream, concurrently with the previous download AND the gather, was slower: 3.99 ms)
            b = k & 1
            if k >= 2:
                st3.stream.wait_event(ev_fetched[b])          # gather buffer b is rewritten by this frame: its download comes first
            d_h0[b].copy_(h0, non_blocking=True)
            d_h0c[b].copy_(h0c, non_blocking=True)
            ts.set_h0(0, d_h0[b].data_ptr(), d_h0c[b].data_ptr())
            g = st3.generate_pipelined(0.016 * k)
            st3.finish()                                      # the user stream waits for this frame's gather
            ev_gathered[b].record(st3.stream)
            d2h.wait_event(ev_gathered[b])
            with torch.cuda.stream(d2h):                      # results leave on a copy stream, under the next step's upload
                out_host.copy_(g[rank], non_blocking=True)
                ev_fetched[b].record(d2h)

        with torch.cuda.stream(st3.stream):
            for k in range(2):
                e2e_step(k)
            d2h.synchronize()
            barrier()
            t0 = time.perf_counter()
            for k in range(Ke):
                e2e_step(2 + k)
            st3.stream.synchronize()
            d2h.synchronize()
            st3.sync()
            e2e_s = max_over_ranks(time.perf_counter() - t0)
        checksum = float(out_host[: N * N].double().abs().sum())
        st3.close()
        e2e = {"value": world * pts_rank * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pts_rank * 16,
               "d2h_bytes_per_step": pts_rank * 28, "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
               "api": "per rank and step: h0/h0conj from pinned host memory (H2D) -> mw_tiles_set_h0 -> mw_tiles_generate_allgather "
                      "-> mw_tiles_wait -> this rank's slot of the GATHERED buffer to pinned host memory (D2H); the all-gather is "
                      "inside the timed region", "height_abs_sum_tile0": checksum}

    # ---- the other BASELINE configs + CPU baseline (rank 0, N = 1 only) ----
    extras = None
    cpu = None
    if rank == 0 and world == 1:
        if not args.no_extra_configs:
            try:
                extras = extra_configs(mw, torch, stream, peak)
            except Exception as e:  # noqa: BLE001
                extras = {"error": repr(e)}
        if not args.no_cpu_baseline:
            from oracle import cref
            cref.build()
            verts = args.cpu_vertices
            rate, dt = cpu_reference_rate(N, 1000, verts, 1)
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{verts} of {N * N} vertices of the same {N}x{N} grid through the literal "
                             f"FFTMesh.Displacement loop (oracle/ref_fftmesh.c), 1 thread as Unity runs it; {dt:.1f} s",
                   "host_threads_available": len(os.sched_getaffinity(0))}
            try:
                frate, fdt = cpu_fft_form_rate(N, 1000)
                cpu["fft_form"] = {"value": frate, "unit": UNIT, "cores": 1, "kind": "port, transform form -- NOT what the reference "
                                   "does (it evaluates the O(N^4) direct sum above)",
                                   "sample": f"one full {N}x{N} frame: numpy fp64 ifft2 x 5 + extraction (oracle/ref_fft64.py), {fdt:.2f} s"}
            except Exception as e:  # noqa: BLE001
                cpu["fft_form"] = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (device Philox4x32-10 + Phillips spectrum, seed 1000+tile)",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu,
        }
        if extras is not None:
            line["configs"] = extras
        if world > 1:
            def arm(r):
                g = r["gather_ms"] / K
                return {"impl": r["impl"], "value": world * pts_rank * K / (r["ms"] * 1e-3), "ms_per_step": r["ms"] / K,
                        "compute_ms_per_step": r["compute_ms"] / K, "allgather_ms_per_step": g,
                        "allgather_busbw_gbs": round(slot_bytes * (world - 1) / (g * 1e-3) / 1e9, 1) if g else None}
            floor_nom = slot_bytes * (world - 1) / (NVLINK_NOMINAL_GBS * 1e9) * 1e3
            floor_meas = slot_bytes * (world - 1) / (NVLINK_PEER_COPY_GBS * 1e9) * 1e3
            line["multi_gpu"] = {
                "compute_only_value": world * pts_rank * K / (compute_ms * 1e-3),
                "allgather_bytes_per_rank": slot_bytes, "ingress_bytes_per_rank_per_step": slot_bytes * (world - 1),
                "ingress_floor_ms": round(floor_nom, 4), "ingress_floor_ms_at_measured_peer_copy": round(floor_meas, 4),
                "ingress_floor_note": f"every rank must RECEIVE (world - 1) x {slot_bytes / 1e6:.1f} MB per step through its NVLink "
                                      f"ingress: {NVLINK_NOMINAL_GBS:.0f} GB/s nominal, {NVLINK_PEER_COPY_GBS:.0f} GB/s measured peer copy",
                "value_ceiling_at_floor": world * pts_rank / (max(floor_nom, compute_ms / K) * 1e-3),
                "arms": {main["impl"]: arm(main), other["impl"]: (arm(other) if "unavailable" not in other else other)},
                "default_arm": main["impl"], "default_arm_rule": "MW_GATHER_AUTO: peer-memory pushes (TMA bulk-copy kernel) at every world size; ncclAllGather if peer memory is unusable",
                **({"default_arm_fallback": f"the peer arm could not be set up on this box ({auto_fallback}); `value` is the ncclAllGather arm's"}
                   if auto_fallback else {}),
                "rank0_numa_binding": numa,
            }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


class _CleanStdout:
    """Library chatter (e.g. "NCCL version ...") must not share stdout with the one JSON line: while active,
    fd 1 points at stderr; emit() writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


_OUT = None


def emit(line):
    if _OUT is not None:
        _OUT.emit(json.dumps(line))
    else:
        print(json.dumps(line), flush=True)


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--resolution", type=int, default=None, help="default: 1024 at N = 1 (configs[2]), 2048 at N > 1 (configs[4])")
    ap.add_argument("--tiles", type=int, default=None, help="tiles per GPU; default: 16 at N = 1, 1 at N > 1")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-vertices", type=int, default=128, help="vertices in the single-thread CPU baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the other BASELINE configs (64^2, 256^2, Gerstner, ...)")
    args = ap.parse_args()
    with _CleanStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_engine(args)


if __name__ == "__main__":
    main()
