#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[2], the one the metric's "disp+Jacobian" is quoted on):
    one step = one frame of a batch of TILES independent 1024 x 1024 Tessendorf grids, each
    h0 -> h(k,t) -> 2-D IFFT -> height + hds + normal + Jacobian whitecap (44 algorithmic B/point),
    outputs left in HBM.  TILES = 16 makes the per-step input (268 MB) and output (470 MB) larger
    than the 126 MB L2, so no flush is needed between timed steps.
With --gpus N > 1 (torchrun, one rank per GPU, NCCL) every rank runs the same batch (weak scaling) and
the step ends with the path's one collective: the in-place all-gather of the final float buffers.

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
ALG_BYTES_KERNEL = {"spectrum_rows": 16 + 24, "cols_extract": 24 + 28}  # + the 24 B/pt intermediate


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world, gather_impl="nccl"):
    N, T = args.resolution, args.tiles
    coll = ("one in-place all-gather of the final float buffers per step, over NVLink peer memory: every rank's copy engines "
            "push its slot into the peers' buffers (CUDA IPC), two 4-byte NCCL all-reduces fence the step"
            if gather_impl == "p2p" else "one in-place all-gather of the final float buffers (NCCL) per step")
    return {
        "workload": f"{T} x ({N}x{N} Tessendorf grid, height+hds+normal+Jacobian whitecap) per GPU per step "
                    f"= BASELINE configs[2] batched",
        "resolution": N, "tiles_per_gpu": T, "points_per_step_per_gpu": T * N * N,
        "outputs": "height,hds,normal,whitecap (28 B/pt)", "algorithmic_bytes_per_point": ALG_BYTES_PIPELINE,
        "l2": f"inputs {T * N * N * 16 / 1e6:.0f} MB + outputs {T * N * N * 28 / 1e6:.0f} MB per step > 126 MB L2; no flush",
        "collective": "none" if world == 1 else coll,
        "parallelism": f"tiles{world}",
    }


# ----------------------------------------------------------------------------- the CUDA engine arm
def run_engine(args):
    import numpy as np
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
    if world > 1:
        # MW_NCCL_HIGH_PRIORITY=1 (experiment, off by default): NCCL's stream gets scheduling priority over the frame kernels,
        # so that the 4-byte fences of the peer-memory gather do not wait for an SM behind a frame
        opts = None
        if os.environ.get("MW_NCCL_HIGH_PRIORITY", "0") == "1":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    N, T, K, W = args.resolution, args.tiles, args.steps, max(args.warmup, 3)
    pts_rank = T * N * N
    peak, peak_src = peaks()

    st = ShardedTiles(N, rank, world, tiles_per_rank=T, base_seed=1000, device=dev)
    ocean, stream = st.ocean, st.stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        # N = 1: the engine's kernels on `stream`.  N > 1: the same, then the path's one collective -- the
        # all-gather of frame k runs on a communication stream under the generation of frame k + 1
        if world > 1:
            st.generate_pipelined(0.016 * k)
        else:
            st.generate_local(0.016 * k)

    with torch.cuda.stream(stream):
        for k in range(W):
            step(k)
        st.finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = mw.native.launch_count()
        with ClockSampler(local) as clk:
            e0.record(stream)
            for k in range(K):
                step(W + k)
            st.finish()          # (N > 1) the last gathers are inside the timed region
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            launches = mw.native.launch_count() - launches0
            # a short timed region gives the sampler too few looks: keep the same load running untimed
            extra = 0
            while len(clk.samples) < 8 and extra < 200:
                for k in range(8):
                    step(k)
                st.finish()
                torch.cuda.synchronize()
                extra += 1
        clocks = clk.summary()
        clocks["sampled"] = "timed region" + (f" + {extra * 8} untimed steps of the same load" if extra else "")
        tms = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())

        # compute-only time (no collective), same loop
        compute_ms = ms
        gather_ms = 0.0
        if world > 1:
            barrier()
            e0.record(stream)
            for k in range(K):
                st.generate_local(0.016 * k)
            e1.record(stream)
            barrier()
            t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            compute_ms = float(t2.item())
            e0.record(stream)
            for k in range(K):
                st.all_gather()
            e1.record(stream)
            barrier()
            t3 = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            gather_ms = float(t3.item())

    value = world * pts_rank * K / (ms * 1e-3)

    # ---- per-kernel durations, live, with CUDA events on the launching stream (MW_PROFILE handle) ----
    roof = None
    if rank == 0:
        prof = mw.Ocean(N, seed=1000, tiles=T, device=local, device_ptrs=True, profile=True)
        prof.set_stream(stream.cuda_stream)
        prof.init_spectrum()
        views = st.slot_views()
        with torch.cuda.stream(stream):
            for k in range(3):
                prof.generate(0.016 * k, views)
            prof.sync()
            prof.kernel_times(reset=True)
            for k in range(K):
                prof.generate(0.016 * k, views)
            kms, kn = prof.kernel_times()
        prof.close()
        names = ["spectrum_rows", "cols_extract"]
        per = {names[i]: kms[i] / max(kn[i], 1) for i in range(2)}              # average duration of ONE launch
        per_step = {names[i]: kms[i] / K for i in range(2)}                      # kernel time per step (all its launches)
        launches_per_step = {names[i]: kn[i] / K for i in range(2)}
        dom = max(per_step, key=per_step.get)
        pts_per_launch = pts_rank / launches_per_step[dom]                       # the engine issues the batch tile group by tile group
        ach = ALG_BYTES_KERNEL[dom] * pts_per_launch / (per[dom] * 1e-3) / 1e9
        frame_ms = compute_ms / K
        roof = {
            "bound": "hbm", "kernel": "k_" + dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
            "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_point": ALG_BYTES_KERNEL[dom], "points_per_launch": int(pts_per_launch),
            "launches_per_step": launches_per_step,
            "avg_launch_ms": {k: round(v, 4) for k, v in per.items()},
            "share_of_step": {k: round(v / sum(per_step.values()), 3) for k, v in per_step.items()},
            "timing": "CUDA events around every launch on the launching stream (MW_PROFILE handle, single stream, no overlap "
                      "between launches), same K steps as the timed region",
            "pipeline": {"algorithmic_bytes_per_point": ALG_BYTES_PIPELINE,
                         "achieved": round(ALG_BYTES_PIPELINE * pts_rank / (frame_ms * 1e-3) / 1e9, 1),
                         "frac": round(ALG_BYTES_PIPELINE * pts_rank / (frame_ms * 1e-3) / 1e9 / peak, 4),
                         "note": "44 B/pt x points per step / whole-step time of the timed region (both kernels, two streams overlapped)"},
        }
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            try:
                t = json.load(open(tr)).get("k_" + dom)
                if t:
                    roof["traffic"] = t["bytes_per_tile"] * pts_per_launch / (N * N)
                    roof["traffic_source"] = t["source"]
            except Exception:  # noqa: BLE001
                pass

    # ---- end to end through the C ABI with HOST buffers (what the C# host calls) ----
    Ke = max(1, min(K, args.e2e_steps))
    # MW_HOST_ASYNC: calls enqueue and return; results leave on the handle's copy stream, so the upload of step k + 1
    # overlaps the download of step k (PCIe is full duplex); everything is complete at host.sync()
    host = mw.Ocean(N, seed=1000 + rank * T, tiles=T, device=local, host_async=True)
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float32).pin_memory()  # noqa: E731
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
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    checksum = float(outs["height"][: N * N].double().abs().sum())
    host.close()
    e2e = {"value": world * pts_rank * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": pts_rank * 16,
           "d2h_bytes_per_step": pts_rank * 28, "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
           "api": "mw_ocean_set_h0 + mw_ocean_generate with pinned host buffers on an MW_HOST_ASYNC handle (upload of step k+1 overlaps download of step k), mw_ocean_sync at the end", "height_abs_sum_tile0": checksum}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
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
            "config": workload_config(args, world, st.gather_impl), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu,
            "achieved_hbm_gbs_pipeline": round(ALG_BYTES_PIPELINE * pts_rank * K / (compute_ms * 1e-3) / 1e9, 1),
        }
        if world > 1:
            slot = st.layout.slot_bytes
            line["multi_gpu"] = {
                "compute_only_value": world * pts_rank * K / (compute_ms * 1e-3), "compute_ms_per_step": compute_ms / K,
                "allgather_ms_per_step": gather_ms / K, "allgather_bytes_per_rank": slot,
                "allgather_busbw_gbs": round(slot * (world - 1) / (gather_ms / K * 1e-3) / 1e9, 1) if gather_ms else None,
                "nvlink_peer_copy_peak_gbs": 770.0, "gather_impl": st.gather_impl, "p2p_error": st.p2p_error,
            }
        emit(line)
    st.close()
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
    ap.add_argument("--resolution", type=int, default=1024)
    ap.add_argument("--tiles", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-vertices", type=int, default=128, help="vertices in the single-thread CPU baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    with _CleanStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            run_engine(args)


if __name__ == "__main__":
    main()
