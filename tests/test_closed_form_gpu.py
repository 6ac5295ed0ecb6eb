"""The CUDA engine, through the C ABI, against closed forms derived by hand from the reference's source (the derivations live next
to the oracle's own closed-form checks: tests/test_oracle.py::_plane_wave_closed_form, tests/test_oracle_renderer.py::
_single_mode_frame) -- no oracle code between the engine and the formula.  Spectra with one to three non-zero entries turn the
whole path (evolution, both FFT passes, extraction, whitecap / the shader chain's stencils) into plane waves with known amplitudes.

Tolerances: fp32 phases (k.x up to ~100 rad, w t up to ~400 rad at t = 60: one ulp of the angle is 3e-5) on amplitudes of O(1)."""
import numpy as np
import pytest

from conftest import max_abs

pytestmark = pytest.mark.gpu

MODES = [(9, 25, 0.3 + 0.1j, -0.2 + 0.15j), (16, 16, 0.5 + 0.0j, 0.1 - 0.1j), (31, 0, -0.15 + 0.2j, 0.05j), (20, 11, 0.35 - 0.2j, 0.0)]


@pytest.mark.parametrize("t", [0.0, 1.7, 60.0])
def test_fftmesh_engine_against_closed_form_plane_waves(mw, t):
    from test_oracle import _plane_wave_closed_form
    N, chop = 32, 0.8
    h0 = np.zeros((N * N, 2), np.float32); hc = np.zeros((N * N, 2), np.float32)
    for n, m, a, b in MODES:
        h0[n * N + m] = (np.real(a), np.imag(a)); hc[n * N + m] = (np.real(b), np.imag(b))
    with mw.Ocean(N, choppiness=chop) as o:
        o.set_h0(h0, hc)
        out = o.generate(t, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
    want = _plane_wave_closed_form(N, float(N), 1.0, chop, MODES, t)
    tol = 2e-4 if t > 10 else 5e-5
    assert max_abs(out["height"][0], want["vertMeow"][:, 1]) < tol
    for k, wk in (("disp", "hds"), ("normal", "normals"), ("vertices", "vertMeow"), ("jacobian", "jacobian")):
        assert max_abs(out[k][0], want[wk]) < tol, (k, max_abs(out[k][0], want[wk]))
    assert max_abs(out["whitecap"][0], want["whitecap"]) < 4 * tol
    assert float(np.abs(want["hds"]).max()) > 0.3 and float(np.ptp(want["whitecap"])) > 0.1   # not a degenerate case


@pytest.mark.parametrize("x0,y0,a,b", [(3, 5, 0.2 - 0.1j, 0.0), (29, 2, 0.0, 0.15 + 0.2j), (6, 27, 0.1 + 0.2j, -0.2 + 0.05j)])
def test_renderer_engine_against_the_single_mode_closed_form(mw, x0, y0, a, b):
    from test_oracle_renderer import _single_mode_frame
    res, length, chop, mult, dt = 4, 40.0, 1.3, 2.0, 0.5
    R = 8 * res
    ini = np.zeros((R, R, 4), np.float32)
    ini[y0, x0] = (a.real, a.imag, np.real(b), np.imag(b))
    a32, b32 = complex(*ini[y0, x0, :2]), complex(*ini[y0, x0, 2:])
    want = _single_mode_frame(res, length, chop, mult, dt, x0, y0, a32, b32)
    with mw.Renderer(res, length, chop, 1.0, (1.0, 0.0), mult, seed1=0.0, seed2=0.0, wrap_repeat=True) as r:
        r.set_initial(ini)
        got = r.generate_texture(dt, names=("displacement", "height", "normal", "white", "jacobian"))
        phase = r.get_phase()[0]
    assert abs(float(phase[y0, x0]) - want["phase_at_mode"]) < 1e-6
    for k in ("displacement", "height"):
        assert max_abs(got[k][0], want[k]) < 2e-6, (k, max_abs(got[k][0], want[k]))
    assert max_abs(got["normal"][0], want["normal"]) < 2e-5
    assert max_abs(got["jacobian"][0, ..., 0], want["jacobian"]) < 2e-6
    assert max_abs(got["white"][0, ..., 0], want["white"]) < 2e-5


def test_gerstner_engine_against_the_shader_formulas(mw):
    """mw_gerstner_from_material / mw_gerstner_append_level_one + mw_gerstner_displace against MistralWaterLib.cginc:71-118 written
    out in float64 (tests/test_oracle_pond.py), with the `* 0.01` of Displacement (:172) applied by from_material."""
    from test_oracle_pond import AMP, DIR_AB, DIR_CD, FREQ, POS, SPEED, STEEP, gerstner4_f64, level_one_f64
    t = 1.7
    g4 = mw.GerstnerWaves.from_material(_Amplitude=AMP / 0.01, _Frequency=FREQ, _Steepness=STEEP, _WSpeed=SPEED, _WDirectionAB=DIR_AB,
                                        _WDirectionCD=DIR_CD)
    out = g4.displace(POS, t)
    assert g4.n_waves == 4
    # out = pos + offsets in fp32 at |pos| <= 40: half an ulp of 40 is 1.9e-6
    assert max_abs(out - POS, gerstner4_f64(POS, t, AMP, FREQ, STEEP, SPEED, DIR_AB, DIR_CD)) < 5e-6
    g5 = mw.GerstnerWaves().append_level_one(0.3, 0.6, 0.9)
    out5 = g5.displace(POS, t)
    assert g5.n_waves == 5
    assert max_abs(out5 - POS, level_one_f64(POS, t, 0.3, 0.6, 0.9)) < 2e-5   # five waves of amplitude ~0.3, fp32 phases up to ~60 rad
