"""Parity against vectors computed by the REAL reference (bindings/DumpGolden.cs run inside Unity on the unmodified
FFTMesh.cs).  No such vectors can be made in the build image, so these tests skip until tests/golden/unity/ (or
$MW_UNITY_VECTORS) holds fftmesh_unity_N*_meta.json files; until then parity is UNPINNED.

When present they pin (a) the CPU oracle -- the literal loop is fed Unity's own h0 / h0conj, so UnityEngine.Random is taken
as given -- and (b) the CUDA engine through mw_ocean_set_h0, both to the tolerance the oracle tests already state."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.environ.get("MW_UNITY_VECTORS", os.path.join(HERE, "golden", "unity"))
METAS = sorted(glob.glob(os.path.join(DIR, "fftmesh_unity_N*_meta.json")))


def _load(meta_path):
    meta = json.load(open(meta_path))
    stem = meta_path[: -len("_meta.json")]
    ld = lambda s: np.load(stem + s).astype(np.float32)  # noqa: E731
    frames = [dict(t=float(t), vertMeow=ld(f"_t{k}_vertMeow.npy"), normals=ld(f"_t{k}_normals.npy"), colors=ld(f"_t{k}_colors.npy"))
              for k, t in enumerate(meta["times"])]
    return meta, ld("_h0.npy"), ld("_h0conj.npy"), ld("_rest.npy"), frames


def test_dump_script_and_fixture_directory_exist():
    """The route to pinned parity is part of the product: the dump script is committed and names what the reader expects."""
    src = open(os.path.join(os.path.dirname(HERE), "bindings", "DumpGolden.cs")).read()
    for piece in ("_h0.npy", "_h0conj.npy", "_rest.npy", "_vertMeow.npy", "_normals.npy", "_colors.npy", "_meta.json",
                  "fftmesh_unity_N", '"EvaluateWaves"', '"GenerateMesh"'):
        assert piece in src, piece
    assert os.path.isdir(os.path.join(HERE, "golden", "unity"))


def test_npy_header_written_by_the_dump_script_is_numpy_readable(tmp_path):
    """WriteNpy's header arithmetic, restated: magic + version + uint16 length + dict padded to a multiple of 64."""
    rows, cols = 256, 3
    d = "{'descr': '<f4', 'fortran_order': False, 'shape': (%d, %d), }" % (rows, cols)
    pad = (64 - (10 + len(d) + 1) % 64) % 64
    header = d + " " * pad + "\n"
    data = np.arange(rows * cols, dtype="<f4")
    p = tmp_path / "x.npy"
    p.write_bytes(b"\x93NUMPY\x01\x00" + len(header).to_bytes(2, "little") + header.encode() + data.tobytes())
    back = np.load(p)
    assert back.shape == (rows, cols) and np.array_equal(back.ravel(), data)


@pytest.mark.skipif(not METAS, reason="no Unity-dumped reference vectors in tests/golden/unity (parity unpinned; see bindings/DumpGolden.cs)")
@pytest.mark.parametrize("meta_path", METAS)
def test_oracle_reproduces_the_reference(cref, meta_path):
    meta, h0, h0c, rest, frames = _load(meta_path)
    N = int(meta["resolution"])
    p = cref.params(N, unit_width=meta["unit_width"], length=meta["length"], choppiness=meta["choppiness"],
                    amplitude=meta["amplitude"], wind=tuple(meta["wind"]))
    verts, _, _ = cref.generate_mesh(p, seed=0)
    assert np.array_equal(verts, rest.reshape(verts.shape))          # FFTMesh.cs:107-112 is exact arithmetic
    for fr in frames:
        lit = cref.evaluate_waves(p, verts, h0.reshape(-1, 2), h0c.reshape(-1, 2), fr["t"], threads=cref.max_threads())
        for k in ("vertMeow", "normals", "colors"):
            scale = max(1.0, float(np.abs(fr[k]).max()))
            assert np.abs(lit[k].reshape(fr[k].shape) - fr[k]).max() <= 2e-5 * scale, (k, fr["t"])


@pytest.mark.gpu
@pytest.mark.skipif(not METAS, reason="no Unity-dumped reference vectors in tests/golden/unity (parity unpinned; see bindings/DumpGolden.cs)")
@pytest.mark.parametrize("meta_path", METAS)
def test_engine_reproduces_the_reference(mw, meta_path):
    meta, h0, h0c, rest, frames = _load(meta_path)
    N = int(meta["resolution"])
    if N < 32:
        pytest.skip("the FFT path starts at N = 32 (N < 32 runs the direct-sum kernel, covered by test_direct_gpu.py)")
    with mw.Ocean(N, unit_width=meta["unit_width"], length=meta["length"], choppiness=meta["choppiness"],
                  amplitude=meta["amplitude"], wind=tuple(meta["wind"])) as o:
        o.set_h0(h0.reshape(1, -1, 2), h0c.reshape(1, -1, 2))
        for fr in frames:
            out = o.generate(fr["t"], names=("vertices", "normal", "colors"))
            for got, k in ((out["vertices"], "vertMeow"), (out["normal"], "normals"), (out["colors"], "colors")):
                scale = max(1.0, float(np.abs(fr[k]).max()))
                assert np.abs(got.reshape(fr[k].shape) - fr[k]).max() <= 1e-4 * scale, (k, fr["t"])
