"""The direct-sum path (csrc/mw_direct_kernels.cuh): grids the transform identity does not cover -- the reference's own
FFT Mesh demo scene (resolution 12, length 12.39: Demo/FFT Mesh.unity:145-152), odd resolutions, power-of-two grids below
32 -- run FFTMesh.Displacement's O(N^2)-per-vertex sum on the GPU.  Checked against the literal C restatement
(oracle/ref_fftmesh.c) at the tolerance the N = 64 literal test uses: the terms are the same fp32 numbers, only the
summation order differs (block tree vs sequential)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENE = dict(resolution=12, unit_width=1.0, length=12.39, choppiness=1.0, amplitude=0.01, wind=(5.0, 3.0))   # FFT Mesh.unity:145-152


def _check(mw, cref, kw, times, seed=1234, tiles=1):
    N = kw["resolution"]
    p = cref.params(N, unit_width=kw["unit_width"], length=kw["length"], choppiness=kw["choppiness"], amplitude=kw["amplitude"],
                    wind=kw["wind"])
    verts, h0, h0c = cref.generate_mesh(p, seed=seed)
    with mw.Ocean(**kw, seed=seed, tiles=tiles) as o:
        assert np.array_equal(o.rest_vertices(), verts)                       # FFTMesh.cs:107-112, odd and even grids
        o.set_h0(np.tile(h0, (tiles, 1, 1)), np.tile(h0c, (tiles, 1, 1)))
        for t in times:
            out = o.generate(t, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
            lit = cref.evaluate_waves(p, verts, h0, h0c, t, threads=cref.max_threads())
            for k in range(tiles):
                for got, want, name in ((out["vertices"][k], lit["vertMeow"], "vertMeow"), (out["normal"][k], lit["normals"], "normals"),
                                        (out["disp"][k], lit["hds"], "hds"), (out["colors"][k], lit["colors"], "colors"),
                                        (out["jacobian"][k, :, 0], lit["jacobian"], "jacobian"),
                                        (out["height"][k, :, 0], lit["vertMeow"][:, 1], "height")):
                    scale = max(1.0, float(np.abs(want).max()))
                    assert np.abs(got.reshape(want.shape) - want).max() <= 2e-5 * scale, (name, t, k)
                assert np.array_equal(out["whitecap"][k, :, 0], out["colors"][k, :, 0])


def test_fft_mesh_demo_scene_runs_through_the_engine(mw, cref):
    """resolution 12, length 12.39: not periodic, not a power of two -- the shipped scene."""
    _check(mw, cref, SCENE, (0.0, 1.7, 60.0))


@pytest.mark.parametrize("N,L", [(16, 16.0), (13, 13.0), (31, 40.5), (50, 1.0)])
def test_small_odd_and_non_periodic_grids(mw, cref, N, L):
    """(50, 1.0) are FFTMesh's own field defaults (FFTMesh.cs:13, :19)."""
    _check(mw, cref, dict(SCENE, resolution=N, length=L), (1.7,))


def test_direct_path_device_init_and_tiles(mw, cref):
    kw = dict(SCENE)
    with mw.Ocean(**kw, seed=77, tiles=3) as o:
        o.init_spectrum()
        h0, hc = o.get_h0()
        p = cref.params(12, length=12.39)
        for k in range(3):
            _, w0, wc = cref.generate_mesh(p, seed=77 + k)
            assert np.allclose(h0[k], w0, rtol=1e-5, atol=1e-8) and np.allclose(hc[k], wc, rtol=1e-5, atol=1e-8)
        out = o.generate(1.7)
        verts, _, _ = cref.generate_mesh(p, seed=77)
        lit = cref.evaluate_waves(p, verts, h0[1], hc[1], 1.7)
        assert np.abs(out["height"][1, :, 0] - lit["vertMeow"][:, 1]).max() <= 2e-5
        assert not np.array_equal(out["height"][0], out["height"][1])


def test_large_grids_still_need_the_periodic_case(mw):
    with pytest.raises(mw.native.MwError) as ei:
        mw.Ocean(512, length=500.0)
    assert ei.value.code == mw.native.MW_E_INVALID_ARG and "periodic" in ei.value.message


def test_host_mirror_runs_the_scene(mw, cref):
    """The MonoBehaviour mirror with the scene's serialized values, two Update() calls."""
    fm = mw.FFTMesh(choppiness=1.0, tDivision=1.0, resolution=12, unitWidth=1.0, length=12.39, wind=(5.0, 3.0), amplitude=0.01, seed=5)
    fm.Awake()
    fm.Update(0.5)
    fm.Update(0.25)
    p = cref.params(12, length=12.39)
    verts, _, _ = cref.generate_mesh(p, seed=5)
    lit = cref.evaluate_waves(p, verts, fm.verttilde, fm.vertConj, 0.75)
    assert abs(fm.timer - 0.75) < 1e-7
    assert np.abs(fm.mesh.vertices - lit["vertMeow"]).max() <= 2e-5 and np.abs(fm.mesh.colors - lit["colors"]).max() <= 2e-5
    fm.close()
