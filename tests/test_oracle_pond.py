"""CPU checks of the pond oracles (oracle/ref_gerstner.c) against forms written independently from the shader source
(MistralWaterLib.cginc), in vectorised float64: the C restatement is what the CUDA kernels are compared with on the GPU, so it is
checked here against the formulas themselves -- signs, which direction component goes to x / z, the order of the `.xz` / `.yw`
swizzles, the `* 0.01` of Displacement, the generalised wave table -- not only against its own frozen outputs."""
import numpy as np
import pytest

from conftest import max_abs

RNG = np.random.default_rng(11)
POS = np.c_[RNG.uniform(-40, 40, 500), RNG.uniform(-1, 1, 500), RNG.uniform(-40, 40, 500)].astype(np.float32)
# the Pond scene's material (Pond Water Mat.mat; tests/test_host_mirror.py checks these numbers against the scene file)
AMP, FREQ, STEEP = 0.05 * 0.01, 0.8, 0.7
SPEED = np.array([1.2, 1.375, 1.1, 1.5], np.float32)
DIR_AB = np.array([0.3, 0.85, 0.85, 0.25], np.float32)
DIR_CD = np.array([0.1, 0.9, 0.5, 0.5], np.float32)


def gerstner4_f64(pos, t, amplitude, frequency, steepness, speed, dAB, dCD):
    """MistralWaterLib.cginc:71-91.  sVertex = worldPos.xzz (:169), so sVertex.xz = (world x, world z)."""
    x, z = pos[:, 0].astype(np.float64), pos[:, 2].astype(np.float64)
    dirs = np.array([[dAB[0], dAB[1]], [dAB[2], dAB[3]], [dCD[0], dCD[1]], [dCD[2], dCD[3]]], np.float64)   # AB.xy, AB.zw, CD.xy, CD.zw
    th = frequency * (dirs[:, 0, None] * x + dirs[:, 1, None] * z) + t * np.asarray(speed, np.float64)[:, None]   # :79-80
    # offs.x = dot(COS, (AB.x, AB.z, CD.x, CD.z)) = sum_w steep * amp * dir_w.x cos; offs.z with the .y components (:85-86)
    ox = (steepness * amplitude * dirs[:, 0, None] * np.cos(th)).sum(0)
    oz = (steepness * amplitude * dirs[:, 1, None] * np.cos(th)).sum(0)
    oy = (amplitude * np.sin(th)).sum(0)                                                                     # :87
    return np.stack([ox, oy, oz], -1)


def level_one_f64(pos, t, amplitude, frequency, steepness):
    """MistralWaterLib.cginc:101-118."""
    amps = [0.7, 0.6, 0.6, 0.7, 0.9]; steeps = [0.95, 0.615, 0.821, 0.462, 0.611]; speeds = [-2.112, 0.6124, -0.878, -3.6234, 1]
    dirs = [(1, -0.2), (-0.9, 1), (0.2, 0.2), (-1.0, 0.77), (0.99, -1.145)]; fs = [0.954, 1.52, 0.44, 0.21, 0.8]
    x, z = pos[:, 0].astype(np.float64), pos[:, 2].astype(np.float64)
    o = np.zeros((len(x), 3))
    for i in range(5):
        th = frequency * fs[i] * (x * dirs[i][0] + z * dirs[i][1]) + speeds[i] * frequency * fs[i] * t
        o[:, 0] += steepness * amplitude * steeps[i] * amps[i] * dirs[i][0] * np.cos(th)
        o[:, 2] += steepness * amplitude * steeps[i] * amps[i] * dirs[i][1] * np.cos(th)
        o[:, 1] += amplitude * amps[i] * np.sin(th)
    return o


@pytest.mark.parametrize("t", [0.0, 1.7, 30.0])
def test_gerstner4_oracle_against_the_shader_formula(cref, t):
    got = cref.gerstner4(POS, t, AMP, FREQ, STEEP, SPEED, DIR_AB, DIR_CD)
    want = gerstner4_f64(POS, t, AMP, FREQ, STEEP, SPEED, DIR_AB, DIR_CD)
    # amplitudes are 5e-4, phases up to ~80 rad in fp32 (ulp 8e-6)
    assert max_abs(got, want) < 2e-8 and np.abs(want).max() > 5e-4   # measured 4e-9


@pytest.mark.parametrize("t", [0.0, 1.7, 30.0])
def test_gerstner_level_one_oracle_against_the_shader_formula(cref, t):
    got = cref.gerstner_level_one(POS, t, 0.3, 0.6, 0.9)
    want = level_one_f64(POS, t, 0.3, 0.6, 0.9)
    assert max_abs(got, want) < 1e-5 and np.abs(want).max() > 0.3   # measured 1.5e-6 (fp32 phases up to ~60 rad)


def test_wave_table_form_contains_both_reference_variants(cref):
    """The 6-float wave table {dir.x, dir.y, freq, rate, amp_xz, amp_y} the engine's kernel runs (and mw_gerstner_from_material /
    _append_level_one fill) reproduces Gerstner and GerstnerLevelOne when filled as oracle/ref_gerstner.c:78-85 says;
    out = pos + offsets (Displacement :176)."""
    t = 2.3
    dirs = [(DIR_AB[0], DIR_AB[1]), (DIR_AB[2], DIR_AB[3]), (DIR_CD[0], DIR_CD[1]), (DIR_CD[2], DIR_CD[3])]
    w4 = [[dx, dy, FREQ, SPEED[k], STEEP * AMP, AMP] for k, (dx, dy) in enumerate(dirs)]
    got = cref.gerstner_table(np.array(w4, np.float32), POS, t)
    assert max_abs(got - POS, gerstner4_f64(POS, t, AMP, FREQ, STEEP, SPEED, DIR_AB, DIR_CD)) < 1e-5   # (pos + offs) - pos in fp32 at |pos| <= 40
    amps = [0.7, 0.6, 0.6, 0.7, 0.9]; steeps = [0.95, 0.615, 0.821, 0.462, 0.611]; speeds = [-2.112, 0.6124, -0.878, -3.6234, 1]
    d5 = [(1, -0.2), (-0.9, 1), (0.2, 0.2), (-1.0, 0.77), (0.99, -1.145)]; fs = [0.954, 1.52, 0.44, 0.21, 0.8]
    A, F, S = 0.3, 0.6, 0.9
    w5 = [[d5[i][0], d5[i][1], F * fs[i], speeds[i] * F * fs[i], S * A * steeps[i] * amps[i], A * amps[i]] for i in range(5)]
    got5, nrm = cref.gerstner_table(np.array(w5, np.float32), POS, t, want_normal=True)
    assert max_abs(got5 - POS, level_one_f64(POS, t, A, F, S)) < 2e-5
    assert np.array_equal(nrm, np.tile(np.float32([0, 1, 0]), (len(POS), 1)))                            # :98, :121


def test_analytic_normal_is_the_normal_of_the_displaced_surface(cref):
    """MW_GERSTNER_NORMAL_ANALYTIC: checked against central differences of the displaced surface itself."""
    t = 0.9
    w = np.array([[0.6, 0.8, 0.5, 1.1, 0.12, 0.2], [-0.3, 0.95, 0.9, -0.7, 0.05, 0.08]], np.float32)
    pos = POS[:64].astype(np.float64)
    n = cref.gerstner_table_normals(w, pos.astype(np.float32), t, mode="analytic").astype(np.float64)

    def surf(p):
        x, z = p[:, 0], p[:, 2]
        o = np.zeros_like(p)
        for dx, dy, f, r, axz, ay in w.astype(np.float64):
            th = f * (dx * x + dy * z) + r * t
            o[:, 0] += axz * dx * np.cos(th); o[:, 2] += axz * dy * np.cos(th); o[:, 1] += ay * np.sin(th)
        return np.stack([x + o[:, 0], o[:, 1], z + o[:, 2]], -1)

    h = 1e-4
    ex, ez = np.array([h, 0, 0]), np.array([0, 0, h])
    dPdx = (surf(pos + ex) - surf(pos - ex)) / (2 * h)
    dPdz = (surf(pos + ez) - surf(pos - ez)) / (2 * h)
    fd = np.cross(dPdz, dPdx)
    fd /= np.linalg.norm(fd, axis=1, keepdims=True)
    assert max_abs(n, fd) < 1e-5 and np.all(n[:, 1] > 0)


def test_wave_mode_oracle_against_the_shader_formula(cref):
    """Wave (:127-152) through Displacement (:160-164) with identity object <-> world matrices."""
    t, amp, freq, s, smooth = 1.3, 10.0, 2.5, 1.3, 0.4
    out, nrm = cref.wave(POS, t, amp, freq, s, smooth)
    p = POS.astype(np.float64)
    a = amp * 0.01
    def y_of(x, z, y0):
        return y0 + np.sin(s * t + x * freq) * a - np.cos(s * t + z * freq) * a
    y0 = y_of(p[:, 0], p[:, 2], p[:, 1]); y1 = y_of(p[:, 0] + 0.05, p[:, 2], p[:, 1]); y2 = y_of(p[:, 0], p[:, 2] + 0.05, p[:, 1])
    y1 = y1 - (y1 - y0) * (1 - smooth); y2 = y2 - (y2 - y0) * (1 - smooth)
    va = np.stack([np.zeros_like(y0), y2 - y0, np.full_like(y0, 0.05)], -1)     # v2 - v0
    vb = np.stack([np.full_like(y0, 0.05), y1 - y0, np.zeros_like(y0)], -1)     # v1 - v0
    c = np.cross(va, vb)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    assert max_abs(out[:, 1], p[:, 1] + y0) < 5e-6 and np.array_equal(out[:, [0, 2]], POS[:, [0, 2]])   # v.vertex.y += offsets.y
    # the normal comes from differences of nearly equal fp32 heights over 0.05 (measured 5e-6)
    assert max_abs(nrm, c) < 5e-5 and np.all(nrm[:, 1] > 0)
