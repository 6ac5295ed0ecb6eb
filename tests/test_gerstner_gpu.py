"""GPU parity tests of the pond Gerstner path against the literal restatement of
MistralWaterLib.cginc (oracle/ref_gerstner.c).

Tolerance: the phase theta is formed with the source's own fp32 roundings on both sides, so the only
difference is sin/cos (MUFU after an exact-ish 3-term reduction, ~5e-7 absolute per wave):
|delta offset| <= 2e-6 * sum_w (|amp_xz| * |dir| + |amp_y|) + 1 ulp of the vertex coordinate.
"""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _tol(tab, pos):
    s = float(np.sum(np.abs(tab[:, 4]) * np.hypot(tab[:, 0], tab[:, 1]) + np.abs(tab[:, 5])))
    return 2e-6 * s + 1.2e-7 * float(np.abs(pos).max())


def test_material_gerstner_matches_literal_4_wave(mw, cref):
    g = golden("gerstner_pond.npz")
    pos, t = g["pos"], float(g["t"])
    gw = mw.GerstnerWaves.from_material(**mw.POND_MATERIAL)
    out = gw.displace(pos, t)
    assert np.abs((out - pos) - g["offsets4"]).max() <= _tol(gw.table(), pos)
    m = mw.POND_MATERIAL
    lit = cref.gerstner4(pos, t, m["_Amplitude"] * 0.01, m["_Frequency"], m["_Steepness"], m["_WSpeed"],
                         m["_WDirectionAB"], m["_WDirectionCD"])
    assert np.abs((out - pos) - lit).max() <= _tol(gw.table(), pos)


def test_level_one_matches_literal_5_wave(mw, cref):
    g = golden("gerstner_pond.npz")
    pos, t = g["pos"], float(g["t"])
    gw = mw.GerstnerWaves().append_level_one(0.1, 2.58, 0.99)
    out = gw.displace(pos, t)
    assert np.abs((out - pos) - g["offsets5"]).max() <= _tol(gw.table(), pos) + 3e-7


def test_single_wave_closed_form(mw):
    gw = mw.GerstnerWaves().append(1.0, 0.0, 0.5, 2.0, 0.3, 0.7)
    x = np.linspace(-20, 20, 4097).astype(np.float32)
    pos = np.stack([x, np.zeros_like(x), 0.3 * x], -1)
    out = gw.displace(pos, 1.25)
    th = 0.5 * x.astype(np.float64) + 2.0 * 1.25
    assert np.abs(out[:, 1] - 0.7 * np.sin(th)).max() <= 3e-6
    assert np.abs(out[:, 0] - (x + 0.3 * np.cos(th))).max() <= 3e-6
    assert np.array_equal(out[:, 2], pos[:, 2])  # dir_y = 0: no z offset


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1024, 1025])
def test_ragged_vertex_counts_and_normals(mw, cref, n):
    gw = mw.pond_wave_table_32()
    rng = np.random.default_rng(n)
    pos = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
    nrm = np.full((n, 3), 7.0, np.float32)
    out = gw.displace(pos, 0.3, normals=nrm)
    ref, rn = cref.gerstner_table(gw.table(), pos, 0.3, want_normal=True) if n else (pos, nrm)
    if n:
        assert np.abs(out - ref).max() <= _tol(gw.table(), pos)
        assert np.array_equal(nrm, rn)  # (0,1,0): MistralWaterLib.cginc:98, :121


def test_config4_32_waves_1m_vertices(mw, cref):
    N = 1024
    ax = ((np.arange(N) - N // 2).astype(np.float32) + np.float32(0.5))
    pos = np.zeros((N * N, 3), np.float32)
    pos[:, 0] = np.repeat(ax, N)
    pos[:, 2] = np.tile(ax, N)
    gw = mw.pond_wave_table_32()
    assert gw.n_waves == 32
    out = gw.displace(pos, 1.7)
    sel = np.random.default_rng(0).choice(N * N, 200_000, replace=False)
    ref = cref.gerstner_table(gw.table(), pos[sel], 1.7)
    assert np.abs(out[sel] - ref).max() <= _tol(gw.table(), pos)
    # size-independent property: linearity in the amplitudes (doubling every amplitude doubles the offset)
    g2 = mw.GerstnerWaves()
    for w in gw.table():
        g2.append(w[0], w[1], w[2], w[3], 2 * w[4], 2 * w[5])
    out2 = g2.displace(pos, 1.7)
    # (out - pos) is only known to an ulp of the coordinate (|pos| <= 512 -> 6.1e-5)
    assert np.abs((out2 - pos) - 2 * (out - pos)).max() <= 3 * 6.2e-5


def test_device_pointer_mode(mw, cref):
    import torch
    gw = mw.pond_wave_table_32(device_ptrs=True)
    pos = torch.rand(4096, 3, device="cuda") * 64 - 32
    out = torch.empty_like(pos)
    gw.displace(pos, 0.9, out=out)
    torch.cuda.synchronize()
    ref = cref.gerstner_table(gw.table(), pos.cpu().numpy(), 0.9)
    assert np.abs(out.cpu().numpy() - ref).max() <= _tol(gw.table(), pos.cpu().numpy())


@pytest.mark.parametrize("n", [1, 7, 1000, 65537])
def test_wave_mode_matches_literal(mw, cref, n):
    """`Wave` displacement (MistralWaterLib.cginc:127-152): vertices <= 2e-6 of the amplitude scale; normals <= 2e-3 --
    they are built from 0.05-unit finite differences of fp32 heights, i.e. quotients of numbers at the rounding level of
    the positions (|x| up to 500 here: ulp 6e-5 against a 0.05 step)."""
    rng = np.random.default_rng(n)
    pos = rng.uniform(-500, 500, (n, 3)).astype(np.float32)
    pos[:, 1] = rng.uniform(-1, 1, n).astype(np.float32)
    want, wn = cref.wave(pos, 12.5, 10.0, 2.58, 1.3, 0.4)
    got, gn = mw.wave_displace(pos, 12.5, 10.0, 2.58, 1.3, 0.4)
    assert np.abs(got - want).max() <= 2e-6 * 1.0 + 1e-6
    assert np.array_equal(got[:, 0], pos[:, 0]) and np.array_equal(got[:, 2], pos[:, 2])
    assert np.abs(gn - wn).max() <= 2e-3 and np.abs(np.linalg.norm(gn, axis=1) - 1).max() <= 1e-5
    only, none = mw.wave_displace(pos, 12.5, 10.0, 2.58, 1.3, 0.4, want_normal=False)
    assert none is None and np.array_equal(only, got)


@pytest.mark.parametrize("n", [1, 7, 4096 + 3])
def test_analytic_and_discarded_normals(mw, cref, n):
    """the normals the reference has code for but overwrites with (0,1,0) (MistralWaterLib.cginc:92-98,
    :122-124).  Tolerance: per-wave sin/cos error 5e-7 times the summed derivative amplitudes, after normalisation."""
    rng = np.random.default_rng(5)
    pos = (rng.uniform(-40, 40, (n, 3))).astype(np.float32)
    gw = mw.pond_wave_table_32()
    tab = gw.table()
    t = 1.7
    nrm = np.empty_like(pos)
    out = gw.displace(pos, t, normals=nrm, normal_mode="analytic")
    want = cref.gerstner_table_normals(tab, pos, t, "analytic")
    amp = float(np.sum((np.abs(tab[:, 4]) + np.abs(tab[:, 5])) * np.abs(tab[:, 2]) * np.hypot(tab[:, 0], tab[:, 1]) ** 2))
    assert np.abs(nrm - want).max() <= 4e-6 * max(1.0, amp)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-6) and (nrm[:, 1] > 0).all()
    assert np.abs(out - cref.gerstner_table(tab, pos, t)).max() <= _tol(tab, pos)      # the displacement is unchanged
    # finite-difference check of what "analytic" means: the normal of the displaced surface itself
    if n == 7:
        h = 1e-3
        p64 = pos.astype(np.float64)
        def surf(p):
            th = tab[:, 2].astype(np.float64) * (p[:, None, 0] * tab[:, 0] + p[:, None, 2] * tab[:, 1]) + tab[:, 3].astype(np.float64) * t
            o = np.stack([(tab[:, 4] * tab[:, 0] * np.cos(th)).sum(1), (tab[:, 5] * np.sin(th)).sum(1), (tab[:, 4] * tab[:, 1] * np.cos(th)).sum(1)], 1)
            return p + o
        dx = (surf(p64 + [h, 0, 0]) - surf(p64 - [h, 0, 0])) / (2 * h)
        dz = (surf(p64 + [0, 0, h]) - surf(p64 - [0, 0, h])) / (2 * h)
        fd = np.cross(dz, dx)
        fd /= np.linalg.norm(fd, axis=1, keepdims=True)
        assert np.abs(fd - want).max() < 1e-4
    # the literal discarded computation of Gerstner() :92-97, with _Smoothing
    nrm2 = np.empty_like(pos)
    gw.displace(pos, t, normals=nrm2, normal_mode="discarded", smoothing=0.35)
    want2 = cref.gerstner_table_normals(tab, pos, t, "discarded", 0.35)
    assert np.abs(nrm2 - want2).max() <= 4e-6
    assert np.all(nrm2[:, 2] == 0.0)
    # and the shipped behaviour stays the default
    nrm3 = np.empty_like(pos)
    gw.displace(pos, t, normals=nrm3)
    assert np.array_equal(nrm3, np.tile(np.float32([0, 1, 0]), (n, 1)))
