"""GPU parity tests of the OceanRenderer path through the C ABI, against
oracle/ref_ocean_renderer.py on the same inputs, the committed golden fixture, and size-independent properties at R = 2048.

Tolerances (floating point path):
  * initial spectrum from the device vs the oracle's (same hash definition): max-abs <= 1e-6 of the image's scale on
    >= 99.9 % of the texels -- the hash frac(sin(x) * 43758.5453) turns a last-bit difference of sin into a different
    random number, so single texels may differ entirely (the reference's own GPU does that everywhere);
  * phase image vs the fp32 oracle: <= 2 ulp of 2 pi per frame;
  * maps vs the fp64 form of the oracle: relative L2 <= 1e-5 (displacement, height, normal; measured 2e-7, 2e-7, 5e-6);
    Jacobian max-abs <= 1e-5 of its scale; white: max-abs <= 1e-4 on 99.99 % of the texels and <= 2e-3 everywhere -- its
    noise term 0.3 |n.xz| inherits the conditioning of OceanNormal's stencil, which normalises a sum of cross products of
    DIFFERENCES of displaced positions: where the choppy surface folds (texel spacing + delta D ~ 0) the vectors shrink
    to the size of fp32 rounding in D and any fp32 evaluation, the reference's included, loses digits there;
  * maps vs the literal fp32 blit chain: relative L2 <= 4e-7 * R -- the chain's own distance from the exact transform
    (unreduced twiddle angles, tests/test_oracle_renderer.py); the CUDA path is closer to the exact transform than that.
"""
import numpy as np
import pytest

from conftest import golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu
F32, F64 = np.float32, np.float64
SCENE = dict(length=434.48, choppiness=0.46, amplitude=0.41, wind=(14.45, 12.0), mult=1.5)  # Demo/Ocean Demo.unity:296-302


@pytest.fixture(scope="module")
def ror():
    from oracle import ref_ocean_renderer
    return ref_ocean_renderer


def _state(ror, res, dtype, initial=None, wrap="clamp", **kw):
    cfg = dict(SCENE)
    cfg.update(kw)
    return ror.RendererState(res, cfg["length"], cfg["choppiness"], cfg["amplitude"], cfg["wind"], 3.7, 8.1, cfg["mult"],
                             dtype, wrap, initial)


def _check_white(got, want):
    d = np.abs(np.asarray(got, F64) - want)
    assert np.quantile(d, 0.9999) <= 1e-4 and d.max() <= 2e-3 and d.mean() <= 2e-6, (np.quantile(d, 0.9999), d.max(), d.mean())


def _engine(mw, res, tiles=1, wrap_repeat=False, **kw):
    cfg = dict(SCENE)
    cfg.update(kw)
    return mw.Renderer(res, cfg["length"], cfg["choppiness"], cfg["amplitude"], cfg["wind"], cfg["mult"], seed1=3.7, seed2=8.1,
                       tiles=tiles, wrap_repeat=wrap_repeat)


@pytest.mark.parametrize("res", [4, 8, 32, 128])
def test_render_initial_matches_the_shader_restatement(mw, ror, res):
    s = _state(ror, res, F32)
    with _engine(mw, res) as r:
        r.render_initial()
        got = r.get_initial()[0]
    scale = float(np.abs(s.initial).max())
    bad = np.abs(got - s.initial).max(-1) > 1e-6 * scale
    assert bad.mean() <= 1e-3, bad.mean()
    assert got[0, 0, 0] == 0.0 and np.all(np.isfinite(got))


@pytest.mark.parametrize("res,frames", [(4, 3), (8, 3), (16, 2), (32, 2)])
def test_frames_vs_fp64_form_and_literal_chain(mw, ror, res, frames):
    s32 = _state(ror, res, F32)
    s64 = _state(ror, res, F64, initial=s32.initial)
    R = 8 * res
    with _engine(mw, res) as r:
        r.set_initial(s32.initial)
        for _ in range(frames):
            got = r.generate_texture(0.016, names=("displacement", "height", "normal", "white", "white_rgba", "jacobian"))
            a, b = s32.generate_texture(0.016), s64.generate_texture(0.016)
        phase = r.get_phase()[0]
    assert max_abs(phase, a["phase"]) <= frames * 2 * 4.8e-7
    for k in ("displacement", "height", "normal"):
        assert rel_l2(got[k][0], b[k]) <= 1e-5, (k, rel_l2(got[k][0], b[k]))
        assert rel_l2(got[k][0], a[k]) <= 4e-7 * R, (k, rel_l2(got[k][0], a[k]))
    _check_white(got["white"][0, ..., 0], b["white"])
    assert max_abs(got["jacobian"][0, ..., 0], b["jacobian"]) <= 1e-5 * max(1.0, float(np.abs(b["jacobian"]).max()))
    w4 = got["white_rgba"][0]
    assert np.array_equal(w4[..., 0], got["white"][0, ..., 0]) and np.array_equal(w4[..., 0], w4[..., 2]) and np.all(w4[..., 3] == 1)
    assert np.all(got["normal"][0, ..., 3] == 1)
    # the height image carries the same complex number twice (SpectrumHeight.shader:46)
    assert np.array_equal(got["height"][0, ..., 0], got["height"][0, ..., 2])


def test_golden_fixture(mw):
    g = golden("renderer_r64.npz")
    with mw.Renderer(int(g["resolution"]), float(g["length"]), float(g["choppiness"]), float(g["amplitude"]), tuple(g["wind"]),
                     float(g["mult"]), seed1=float(g["seed1"]), seed2=float(g["seed2"])) as r:
        r.set_initial(g["initial"])
        for _ in range(int(g["frames"])):
            m = r.generate_texture(float(g["dt"]))
        for k in ("displacement", "height", "normal"):
            assert rel_l2(m[k][0], g[k]) <= 4e-7 * 64, k
        assert max_abs(m["white"][0, ..., 0], g["white"]) <= 2e-4
        assert max_abs(r.get_phase()[0], g["phase"]) <= 3e-6


def test_wrap_repeat_flag(mw, ror):
    s = _state(ror, 8, F64, wrap="repeat")
    with _engine(mw, 8, wrap_repeat=True) as r:
        r.set_initial(s.initial)
        got = r.generate_texture(0.25)
    b = s.generate_texture(0.25)
    assert rel_l2(got["normal"][0], b["normal"]) <= 1e-5
    _check_white(got["white"][0, ..., 0], b["white"])


def test_ocean_demo_scene_component_lifecycle(mw, ror):
    """The MonoBehaviour mirror with the Ocean Demo scene's serialized values (resolution 128 -> 1024^2 maps)."""
    c = mw.OceanRenderer(mult=1.5, resolution=128, length=434.48, choppiness=0.46, amplitude=0.41, wind=(14.45, 12.0),
                         randomSeed1=3.7, randomSeed2=8.1)
    c.Awake()
    assert c.mesh.vertices.shape == (128 * 128, 3) and c.mesh.indices.shape == (127 * 127 * 6,)
    ini = c.initialTexture
    s = ror.RendererState(128, 434.48, 0.46, 0.41, (14.45, 12.0), 3.7, 8.1, 1.5, F64, initial=ini)
    for _ in range(2):
        c.Update(0.02)
        b = s.generate_texture(0.02)
    assert c.displacementTexture.shape == (1024, 1024, 4) and c.whiteTexture.shape == (1024, 1024)
    for got, k in ((c.displacementTexture, "displacement"), (c.heightTexture, "height"), (c.normalTexture, "normal")):
        assert rel_l2(got, b[k]) <= 1e-5, (k, rel_l2(got, b[k]))
    _check_white(c.whiteTexture, b["white"])
    # Update's parameter refresh (:98-109): a new wind re-renders the initial spectrum, a new choppiness does not
    c.choppiness = 0.9
    c.Update(0.02)
    assert np.array_equal(c.initialTexture, ini)
    c.wind = (3.0, -4.0)
    c.Update(0.02)
    assert not np.array_equal(c.initialTexture, ini)
    c.close()


def test_default_resolution_2048_maps_against_the_fp64_form(mw, ror):
    """R = 2048 (OceanRenderer's default resolution 256, OceanRenderer.cs:13) against the oracle itself, not only properties:
    two frames of all four maps vs the fp64 evaluation of the shader chain on the same initial spectrum (~30 s of numpy)."""
    res = 256
    with _engine(mw, res) as r:
        r.render_initial()
        ini = r.get_initial()
        s = _state(ror, res, F64, initial=ini[0])
        for _ in range(2):
            got = r.generate_texture(0.02)
            want = s.generate_texture(0.02)
        assert max_abs(r.get_phase()[0], want["phase"]) <= 3e-6
    for k in ("displacement", "height"):
        assert rel_l2(got[k][0], want[k]) <= 1e-5, (k, rel_l2(got[k][0], want[k]))
    # OceanNormal's stencil normalises cross products of DIFFERENCES of displaced positions one texel apart: its conditioning
    # grows as the texel shrinks (434.48 / 2048 = 0.21 here against 0.42 at R = 1024, where 5e-6 is measured): 1.13e-5 measured
    assert rel_l2(got["normal"][0], want["normal"]) <= 2.5e-5, rel_l2(got["normal"][0], want["normal"])
    _check_white(got["white"][0, ..., 0], want["white"])


def test_full_size_properties_2048(mw):
    """R = 2048 (the default resolution 256, OceanRenderer.cs:13): properties that need no oracle run.
    Linearity of the whole spectral chain in the initial spectrum; phase accumulation; unit normals."""
    res = 256
    rng = np.random.default_rng(5)
    with _engine(mw, res) as r:
        r.render_initial()
        ini = r.get_initial()
        m1 = r.generate_texture(0.016, names=("displacement", "height", "normal", "white"))
        ph1 = r.get_phase()
        # same phase, doubled spectrum -> doubled displacement / height
        r.set_initial(2.0 * ini)
        r.set_phase(np.zeros_like(ph1))
        m2 = r.generate_texture(0.016, names=("displacement", "height"))
        assert np.array_equal(r.get_phase(), ph1)
        assert rel_l2(m2["displacement"], 2.0 * m1["displacement"]) <= 1e-6
        assert rel_l2(m2["height"], 2.0 * m1["height"]) <= 1e-6
        n = m1["normal"][0, ..., :3].astype(np.float64)
        assert np.abs(np.linalg.norm(n, axis=-1) - 1).max() <= 1e-5
        assert m1["white"].min() >= 0.0 and m1["white"].max() <= 1.0
        # the mean of the height image is the k = 0 input texel: h0 is 0 there (Phillips(0) = 0) but h0conj is not --
        # InitialSpectrum.shader:47 evaluates it at index R - 1 (the mirror-index quirk) -- and its phase rate is 0
        want_mean = float(ini[0, 0, 0, 0]) + float(ini[0, 0, 0, 2])
        assert abs(m1["height"][0, ..., 0].astype(np.float64).mean() - want_mean) <= 1e-6 * float(np.abs(m1["height"]).max())
        # spot check a few texels of the height image against the direct DFT definition
        h0 = ini[0].astype(np.float64)
        phase = ph1[0].astype(np.float64)
        h = (h0[..., 0] + 1j * h0[..., 1]) * np.exp(1j * phase) + (h0[..., 2] + 1j * h0[..., 3]) * np.exp(-1j * phase)
        R = 8 * res
        for (y, x) in ((0, 0), (5, 1999), (1024, 1024), (2047, 3)):
            e = np.exp(-2j * np.pi * (np.arange(R) * y)[:, None] / R) * np.exp(-2j * np.pi * (np.arange(R) * x)[None, :] / R)
            want = (h * e).sum()
            got = m1["height"][0, y, x, 0] + 1j * m1["height"][0, y, x, 1]
            assert abs(got - want) <= 2e-5 * float(np.abs(m1["height"]).max())


def test_tiles_and_device_pointers(mw):
    import torch
    res = 16
    with _engine(mw, res, tiles=3) as r:
        r.render_initial()
        ini = r.get_initial()
        m = r.generate_texture(0.03)
    assert not np.array_equal(ini[0], ini[1])  # tile t uses seeds + t
    for t in range(3):
        with mw.Renderer(res, SCENE["length"], SCENE["choppiness"], SCENE["amplitude"], SCENE["wind"], SCENE["mult"]) as one:
            one.set_initial(ini[t:t + 1])
            m1 = one.generate_texture(0.03)
        for k in m:
            assert np.array_equal(m[k][t], m1[k][0]), (t, k)
    R = 8 * res
    with mw.Renderer(res, SCENE["length"], SCENE["choppiness"], SCENE["amplitude"], SCENE["wind"], SCENE["mult"], tiles=3,
                     device_ptrs=True) as d:
        d.set_initial(torch.from_numpy(ini).cuda())
        bufs = {"displacement": torch.empty(3, R, R, 4, device="cuda"), "height": torch.empty(3, R, R, 4, device="cuda"),
                "normal": torch.empty(3, R, R, 4, device="cuda"), "white": torch.empty(3, R, R, device="cuda")}
        d.generate_texture(0.03, bufs)
        d.sync()
        for k in ("displacement", "height", "normal"):
            assert np.array_equal(bufs[k].cpu().numpy(), m[k])
        assert np.array_equal(bufs["white"].cpu().numpy(), m["white"][..., 0])


def test_errors(mw):
    with pytest.raises(mw.native.MwError) as e:
        mw.Renderer(12, 100.0)
    assert e.value.code == mw.native.MW_E_INVALID_ARG
    with pytest.raises(mw.native.MwError):
        mw.Renderer(512, 100.0)
    with mw.Renderer(4, 32.0) as r:
        with pytest.raises(mw.native.MwError) as e:
            r.generate_texture(0.1)
        assert e.value.code == mw.native.MW_E_STATE


@pytest.mark.parametrize("n", [2, 5, 64, 256])
def test_mesh_generate_matches_the_loop(mw, ror, n):
    m = mw.generate_mesh(n, 1.25)
    v, nr, uv, idx = ror.generate_mesh(n, 1.25)
    assert np.array_equal(m.vertices, v) and np.array_equal(m.normals, nr) and np.array_equal(m.indices, idx)
    assert np.array_equal(m.uv, uv)
