"""CPU tests of the host-side mirror of the reference's C# / shader parameter surface."""
import dataclasses
import math
import os
import re

import numpy as np

REF = "/root/reference/Assets/Mistral Water"


def test_fftmesh_fields_mirror_the_monobehaviour(mw):
    f = {x.name: x.default for x in dataclasses.fields(mw.FFTMesh)}
    # FFTMesh.cs:9-23, names and defaults
    want = dict(choppiness=1.0, tDivision=1.0, resolution=50, unitWidth=1.0, generate=False, length=1.0,
                wind=(1.0, 1.0), amplitude=1.0)
    for k, v in want.items():
        assert f[k] == v, k
    for m in ("Awake", "Update", "SetParams", "GenerateMesh", "EvaluateWaves"):
        assert callable(getattr(mw.FFTMesh, m))
    src = os.path.join(REF, "Scripts", "FFTMesh.cs")
    if os.path.exists(src):  # only in the build container; never on the GPU box
        txt = open(src, encoding="utf-8-sig").read()
        public = re.findall(r"public\s+(?:float|int|bool|Vector2)\s+(\w+)", txt)
        assert set(public) == set(want)


def test_pond_material_table_matches_the_shader_formulas(mw):
    m = mw.POND_MATERIAL
    g = mw.GerstnerWaves.from_material(**m)
    tab = g.table()
    assert tab.shape == (4, 6)
    amp = np.float32(m["_Amplitude"]) * np.float32(0.01)  # MistralWaterLib.cginc:172
    dirs = [m["_WDirectionAB"][0:2], m["_WDirectionAB"][2:4], m["_WDirectionCD"][0:2], m["_WDirectionCD"][2:4]]
    for k in range(4):
        assert np.allclose(tab[k, 0:2], dirs[k])
        assert tab[k, 2] == np.float32(m["_Frequency"]) and tab[k, 3] == np.float32(m["_WSpeed"][k])
        assert tab[k, 4] == np.float32(m["_Steepness"]) * amp and tab[k, 5] == amp
    g.append_level_one(0.1, 2.58, 0.99)
    t5 = g.table()[4:]
    fs = [0.954, 1.52, 0.44, 0.21, 0.8]
    speeds = [-2.112, 0.6124, -0.878, -3.6234, 1.0]
    amps = [0.7, 0.6, 0.6, 0.7, 0.9]
    for i in range(5):
        assert math.isclose(t5[i, 2], 2.58 * fs[i], rel_tol=1e-6)
        assert math.isclose(t5[i, 3], speeds[i] * 2.58 * fs[i], rel_tol=1e-6)
        assert math.isclose(t5[i, 5], 0.1 * amps[i], rel_tol=1e-6)


def test_pond_material_values_match_the_scene_file(mw):
    mat = os.path.join(REF, "Materials", "Pond Water Mat.mat")
    if not os.path.exists(mat):
        return
    txt = open(mat).read()
    for key in ("_Amplitude", "_Frequency", "_Steepness"):
        assert float(re.search(rf"- {key}: ([-\d.]+)", txt).group(1)) == mw.POND_MATERIAL[key]
    for key in ("_WDirectionAB", "_WDirectionCD", "_WSpeed"):
        m = re.search(rf"- {key}: \{{r: ([-\d.]+), g: ([-\d.]+), b: ([-\d.]+), a: ([-\d.]+)\}}", txt)
        assert tuple(float(x) for x in m.groups()) == tuple(mw.POND_MATERIAL[key])


def test_wave_table_32_is_deterministic_and_bounded(mw):
    a, b = mw.pond_wave_table_32().table(), mw.pond_wave_table_32().table()
    assert a.shape == (32, 6) and np.array_equal(a, b)
    assert np.abs(a[:, 0:2]).max() <= 1.2 + 1e-6


def test_tile_layout_and_wind_rotation(mw):
    from mistral_water_b200.tiles import FLOATS_PER_POINT, TileLayout, tile_wind
    L = TileLayout(2048, 8, 1)
    assert FLOATS_PER_POINT == 7 and L.slot_bytes == 2048 * 2048 * 28  # 117.4 MB per rank (config 5)
    assert L.field_range("height") == (0, 2048 * 2048)
    assert L.field_range("whitecap")[1] == L.slot_floats
    L2 = TileLayout(64, 4, 3)
    assert [L2.owner(t) for t in (0, 2, 3, 11)] == [(0, 0), (0, 2), (1, 0), (3, 2)]
    w = tile_wind((5.0, 3.0), 2)  # 90 degrees
    assert np.allclose(w, (-3.0, 5.0))
    assert np.allclose(np.hypot(*tile_wind((5.0, 3.0), 5)), np.hypot(5.0, 3.0))


def test_bench_helpers_run_without_a_gpu(cref):
    """bench.py's CPU-side pieces: the transform-form CPU figure and the workload description of both gather arms."""
    import argparse
    import bench

    rate, dt = bench.cpu_fft_form_rate(64, 1234)
    assert rate > 0 and dt > 0
    args = argparse.Namespace(resolution=1024, tiles=16)
    one = bench.workload_config(args, 1)
    assert one["collective"] == "none" and one["algorithmic_bytes_per_point"] == 44 and one["parallelism"] == "tiles1"
    many = argparse.Namespace(resolution=None, tiles=None)
    bench.defaults_for_world(many, 8)
    assert (many.resolution, many.tiles) == (2048, 1)            # N > 1 runs BASELINE configs[4]
    cfg = bench.workload_config(many, 8)
    assert "ncclAllGather" in cfg["collective"] and "configs[4]" in cfg["workload"] and "117.4 MB" in cfg["workload"]
    solo = argparse.Namespace(resolution=None, tiles=None)
    bench.defaults_for_world(solo, 1)
    assert (solo.resolution, solo.tiles) == (1024, 16)


def test_bench_has_no_collective_under_a_rank_local_condition():
    """Round-1 defect: a clock-sampler top-up loop with a rank-local trip count contained collectives and desynchronised
    the ranks.  The loop that remains may only call generate_local (no collective inside)."""
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    m = re.search(r"while len\(clk\.samples\) < 8 and extra < 200:(.*?)extra \+= 1", src, re.S)
    assert m, "top-up loop not found"
    body = m.group(1)
    assert "generate_local" in body
    for forbidden in ("generate_pipelined", "all_gather", "all_reduce", "barrier(", "generate("):
        assert forbidden not in body, forbidden
