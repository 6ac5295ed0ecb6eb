"""CPU tests of the OceanRenderer-path oracle (oracle/ref_ocean_renderer.py): internal consistency of the literal fp32
blit chain against its fp64 transform form, known answers of each shader restatement, and the golden fixture."""
import numpy as np
import pytest

from conftest import golden, rel_l2
from oracle import ref_ocean_renderer as R

F32, F64 = np.float32, np.float64
SCENE = dict(length=434.48, choppiness=0.46, amplitude=0.41, wind=(14.45, 12.0), mult=1.5)  # Demo/Ocean Demo.unity:296-302


def test_get_wave_is_fft_ordered():
    n = np.arange(8, dtype=F32) + F32(0.5)
    kx, kz = R.get_wave(n, n, 16.0, 8, F64)
    want = 2 * 3.1415926536 * np.array([0, 1, 2, 3, -4, -3, -2, -1]) / 16.0
    assert np.allclose(kx, want, rtol=1e-12) and np.allclose(kz, want, rtol=1e-12)


def test_phillips_known_answers():
    n = np.array([0.5], F32)
    assert R.phillips(n, n, 1e-4, (10.0, 0.0), 64, 64.0, F64)[0] == 0.0          # |k| < EPSILON -> 0 (FFTCommon.cginc:75)
    # a wave along the wind carries energy, one across it none: (k.w)^2
    a = R.phillips(np.array([3.5], F32), np.array([0.5], F32), 1e-4, (10.0, 0.0), 64, 64.0, F64)[0]
    b = R.phillips(np.array([0.5], F32), np.array([3.5], F32), 1e-4, (10.0, 0.0), 64, 64.0, F64)[0]
    assert a > 0 and b == 0.0
    # closed form, damping 0.01 (:82)
    k = 2 * 3.1415926536 * 3 / 64.0
    l = 100.0 / 9.81
    want = 1e-4 * np.exp(-1 / (k * k * l * l)) / k ** 4 * np.exp(-k * k * l * l * 1e-4)
    assert abs(a - want) <= 1e-9 * want * 1e3


def test_initial_spectrum_mirror_index_quirk():
    """InitialSpectrum.shader:47 calls Phillips(R - n, R - m) with n = x + .5: GetWave then sees index R - 1 - x, so the
    'conjugate' spectrum of texel x uses the wave vector of texel R - 1 - x (not R - x)."""
    Rr = 32
    img = R.initial_spectrum(Rr, 32.0, 1e-4, (10.0, 3.0), 2.0, 5.0, F64)
    u, v = R.texcoords(Rr, F64)
    phi_at = lambda ix, iy: R.phillips(np.array([ix + 0.5]), np.array([iy + 0.5]), 1e-4, (10.0, 3.0), Rr, 32.0, F64)[0]
    x, y = 5, 9
    mag2 = img[y, x, 2] ** 2 + img[y, x, 3] ** 2
    bx, by = R.htilde0(u[y:y + 1, x:x + 1], v[y:y + 1, x:x + 1], F64(2.0), F64(5.0), np.array([[phi_at(Rr - 1 - x, Rr - 1 - y)]]), F64)
    assert np.isclose(mag2, bx[0, 0] ** 2 + by[0, 0] ** 2, rtol=1e-12)
    assert np.all(np.isfinite(img)) and img[0, 0, 0] == 0.0  # texel (0, 0): k = 0 -> phi1 = 0


def test_hash_noise_is_clamped_and_deterministic():
    u, v = R.texcoords(64, F32)
    r = R.uv_random(u, v, 10.612, 0.75, F32)
    assert r.min() >= 0.0 and r.max() < 1.0 and r.std() > 0.2
    assert np.array_equal(r, R.uv_random(u, v, 10.612, 0.75, F32))


def test_dispersion_accumulates_modulo_two_pi():
    Rr = 32
    ph = np.zeros((Rr, Rr), F32)
    rate = R.dispersion_rate(Rr, 32.0, F32)
    assert rate[0, 0] == 0.0 and np.isclose(rate[0, 1], np.sqrt(9.81 * (2 * 3.1415926536 / 32.0) * (1 + (2 * 3.1415926536 / 32) ** 2 / 370 ** 2)), rtol=1e-6)
    for _ in range(200):
        ph = R.dispersion_step(ph, Rr, 32.0, 0.05, F32)
    assert ph.min() >= 0.0 and ph.max() < 2 * 3.1415927
    exact = np.mod(200 * 0.05 * rate.astype(F64), 2 * 3.1415926536)
    d = np.abs(ph - exact)
    assert np.minimum(d, 2 * 3.1415926536 - d).max() < 2e-4   # fp32 accumulation over 200 frames


@pytest.mark.parametrize("Rr", [32, 64, 128])
def test_stockham_chain_is_the_forward_dft(Rr):
    rng = np.random.default_rng(Rr)
    x = rng.standard_normal((Rr, Rr, 4))
    want = R.transform(x, F64)
    assert rel_l2(R.stockham_chain(x, F64), want) < 2e-11 * Rr  # exact up to the shader's 11-digit PI (angles up to pi R)
    # the literal fp32 chain: its twiddle angles -2 PI index / S are NOT reduced, so they carry ulp(pi R) of error
    assert rel_l2(R.stockham_chain(x.astype(F32), F32), want) < 4e-7 * Rr


def test_frame_literal_fp32_vs_fp64_form():
    """The literal fp32 blit chain must sit inside the band the CUDA path is held to around the fp64 form -- scaled by
    the chain's own unreduced-twiddle error (previous test)."""
    s32 = R.RendererState(8, dtype=F32, seed1=3.7, seed2=8.1, **SCENE)
    s64 = R.RendererState(8, dtype=F64, seed1=3.7, seed2=8.1, initial=s32.initial, **SCENE)
    for _ in range(3):
        a, b = s32.generate_texture(0.016), s64.generate_texture(0.016)
    for k in ("displacement", "height", "normal", "white"):
        assert rel_l2(a[k], b[k]) < 1e-4, k
    assert rel_l2(a["phase"], b["phase"]) < 1e-6


def test_normal_and_whitecap_of_a_flat_sea():
    Rr = 32
    z = np.zeros((Rr, Rr, 4))
    n = R.ocean_normal(z, z, Rr, 32.0, "clamp", F64)
    assert np.allclose(n[..., :3], [0, 1, 0]) and np.all(n[..., 3] == 1)
    w, jac = R.white_cap(z, n, Rr, 4, "clamp", F64)
    assert np.allclose(jac, 1) and np.allclose(w, 0)


def test_clamp_and_repeat_differ_only_near_the_border():
    s = R.RendererState(4, dtype=F64, seed1=1.0, seed2=2.0, **SCENE)
    m = s.generate_texture(0.5)
    nr = R.ocean_normal(m["displacement"], m["height"], s.R, s.length, "repeat", F64)
    assert np.array_equal(nr[1:-1, 1:-1], m["normal"][1:-1, 1:-1]) and not np.array_equal(nr, m["normal"])
    wr, _ = R.white_cap(m["displacement"], nr, s.R, 4, "repeat", F64)
    assert np.array_equal(wr[8:-8, 8:-8], m["white"][8:-8, 8:-8])


def test_generate_mesh_matches_the_loop():
    v, n, uv, idx = R.generate_mesh(5, 2.0)
    assert idx.size == 4 * 4 * 6 and idx.max() == 24 and idx.min() == 0
    assert np.allclose(v[0], [-4.0, 0, -4.0]) and np.allclose(v[-1], [4.0, 0, 4.0])       # odd resolution: no half offset
    v2, _, uv2, _ = R.generate_mesh(4, 1.0)
    assert np.allclose(v2[0], [-1.5, 0, -1.5]) and np.allclose(uv2[-1], [1, 1]) and np.allclose(uv2[1], [0, 1 / 3])
    assert list(idx[:6]) == [0, 1, 5, 1, 2, 6]


def test_golden_renderer_fixture():
    g = golden("renderer_r64.npz")
    s = R.RendererState(int(g["resolution"]), float(g["length"]), float(g["choppiness"]), float(g["amplitude"]),
                        tuple(g["wind"]), float(g["seed1"]), float(g["seed2"]), float(g["mult"]), F32)
    assert np.array_equal(s.initial, g["initial"])
    for f in range(int(g["frames"])):
        m = s.generate_texture(float(g["dt"]))
    for k in ("displacement", "height", "normal", "white", "phase"):
        assert np.allclose(m[k], g[k], rtol=0, atol=1e-6 * max(1.0, float(np.abs(g[k]).max()))), k


# ---------------------------------------------------------------- a closed form derived from the shader sources, not from the oracle
def _single_mode_frame(res, length, chop, mult, dt, x0, y0, a, b):
    """OceanRenderer.GenerateTexture for an initial image with ONE non-zero texel (x0, y0) = (h0, h0conj) = (a, b), first frame
    (phase images start black), repeat wrap, all in float64 and in closed form:
      Dispersion.shader:37-40 + FFTCommon.cginc:101-114   phi = fmod(sqrt(G |k| (1 + |k|^2 / 370^2)) dt mult, 2 PI)
      Spectrum.shader:45-50                               h = a e^{i phi} + b e^{-i phi};  hx = -i h kx / w chop;  hz = -i h kz / w chop
      Stockham.shader + OceanRenderer.cs:229-298          forward-sign, un-normalised 2-D DFT: one texel -> a plane wave
      OceanNormal.shader:32-56, WhiteCap.shader:33-45     stencils of that plane wave (taps 1 texel / 8 texels apart)."""
    R_ = 8 * res
    PI, G = 3.1415926536, 9.81
    nx = x0 if x0 < R_ / 2 else x0 - R_                      # GetWave, FFTCommon.cginc:58-67 (after its n -= 0.5)
    ny = y0 if y0 < R_ / 2 else y0 - R_
    kx, kz = 2 * PI * nx / length, 2 * PI * ny / length
    kl = np.hypot(kx, kz)
    phi = np.fmod(np.sqrt(G * kl * (1 + kl * kl / 370 / 370)) * dt * mult, 2 * PI)
    h = a * np.exp(1j * phi) + b * np.exp(-1j * phi)
    w = max(1e-4, kl)
    hx, hz = -1j * h * kx / w * chop, -1j * h * kz / w * chop
    Y, X = np.meshgrid(np.arange(R_), np.arange(R_), indexing="ij")      # images are [y][x]

    def plane(c, dx=0, dy=0):
        return c * np.exp(-2j * np.pi * (x0 * (X + dx) + y0 * (Y + dy)) / R_)

    Dx, Dz, H = plane(hx), plane(hz), plane(h)
    disp = np.stack([Dx.real, Dx.imag, Dz.real, Dz.imag], -1)
    height = np.stack([H.real, H.imag, H.real, H.imag], -1)
    ts = length / R_
    centre = np.stack([Dx.real, Dx.imag, Dz.real], -1)                  # disp.rgb "as written" (:44)

    def vec(dx, dy):                                                    # GetVec :32-37 = (disp.r, height.r, disp.b)
        return np.stack([plane(hx, dx, dy).real, plane(h, dx, dy).real, plane(hz, dx, dy).real], -1)

    right = np.array([ts, 0, 0]) + vec(1, 0) - centre
    left = np.array([-ts, 0, 0]) + vec(-1, 0) - centre
    top = np.array([0, 0, -ts]) + vec(0, -1) - centre
    bottom = np.array([0, 0, ts]) + vec(0, 1) - centre
    s = np.cross(right, top) + np.cross(top, left) + np.cross(left, bottom) + np.cross(bottom, right)
    n = s / np.linalg.norm(s, axis=-1, keepdims=True)
    st = R_ // res                                                       # 1 / _Length in uv, _Length = resolution (:306) -> 8 texels

    def rb(dx, dy):
        return np.stack([plane(hx, dx, dy).real, plane(hz, dx, dy).real], -1)

    ddy = -0.5 * (rb(0, -st) - rb(0, st)) / 8
    ddx = -0.5 * (rb(-st, 0) - rb(st, 0)) / 8
    jac = (1 + ddx[..., 0]) * (1 + ddy[..., 1]) - ddx[..., 1] * ddy[..., 0]
    turb = np.maximum(0, 1 - jac + 0.3 * np.hypot(n[..., 0], n[..., 2]))
    c = np.clip(turb, 0, 1)
    return {"displacement": disp, "height": height, "normal": np.concatenate([n, np.ones_like(n[..., :1])], -1),
            "white": c * c * (3 - 2 * c), "jacobian": jac, "phase_at_mode": phi}


@pytest.mark.parametrize("x0,y0,a,b", [(3, 5, 0.2 - 0.1j, 0.0), (29, 2, 0.0, 0.15 + 0.2j), (6, 27, 0.1 + 0.2j, -0.2 + 0.05j)],
                         ids=["h0", "h0conj-negative-kx", "both-negative-kz"])
def test_single_mode_frame_against_the_closed_form(x0, y0, a, b):
    """The whole chain -- phase step, spectra, 2 x log2(R) Stockham blits per axis, normal and whitecap stencils -- for a spectrum
    with one non-zero texel, against the plane wave it must produce.  Catches a wrong DFT sign, a transposed image, a mis-packed
    channel, the sign of -MultByI, the FFT-ordered wave vector and the tap distances, none of which oracle-vs-oracle fixtures see."""
    res, length, chop, mult, dt = 4, 40.0, 1.3, 2.0, 0.5      # dt * mult is formed in fp32 (OceanRenderer.cs:223): keep it exact
    Rr = 8 * res
    ini = np.zeros((Rr, Rr, 4), F32)
    ini[y0, x0] = (a.real, a.imag, np.real(b), np.imag(b))
    a32, b32 = complex(*ini[y0, x0, :2]), complex(*ini[y0, x0, 2:])       # what the image really holds (fp32-rounded)
    want = _single_mode_frame(res, length, chop, mult, dt, x0, y0, a32, b32)
    st = R.RendererState(res, length, chop, 1.0, (1.0, 0.0), 0.0, 0.0, mult=mult, dtype=F64, wrap="repeat", initial=ini)
    got = st.generate_texture(dt)
    assert abs(got["phase"][y0, x0] - want["phase_at_mode"]) < 1e-12
    for k in ("displacement", "height", "normal", "white", "jacobian"):
        assert np.abs(got[k] - want[k]).max() < 2e-9, (k, np.abs(got[k] - want[k]).max())   # the shader's 11-digit PI in the twiddles
    assert np.abs(want["displacement"]).max() > 5e-2 and np.ptp(want["white"]) > 1e-3           # the case is not degenerate
