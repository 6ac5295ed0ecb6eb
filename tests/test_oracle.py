"""CPU tests of the oracle itself: the literal C restatement of FFTMesh.cs against (a) the known-answer
properties derivable from the reference source, (b) the committed golden
fixtures, (c) the fp64 transform form used for large grids, (d) the Stockham.shader restatement."""
import numpy as np
import pytest

from conftest import golden, max_abs, rel_l2


# ---------------------------------------------------------------- RNG stand-in
def test_philox_known_answers(cref):
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors)."""
    import ctypes as C
    # ref_uniforms exposes (seed, idx) only; check the raw generator through a tiny C shim instead
    import subprocess, tempfile, os, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "ref_philox.h"
        int main(void){ uint32_t o[4];
          ref_philox4x32_10(0,0,0,0,0,0,o); printf("%08x %08x %08x %08x\\n",o[0],o[1],o[2],o[3]);
          ref_philox4x32_10(0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,o); printf("%08x %08x %08x %08x\\n",o[0],o[1],o[2],o[3]);
          ref_philox4x32_10(0x243f6a88u,0x85a308d3u,0x13198a2eu,0x03707344u,0xa4093822u,0x299f31d0u,o); printf("%08x %08x %08x %08x\\n",o[0],o[1],o[2],o[3]);
          return 0; }""")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "kat.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "kat")
        subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O1", "-I", os.path.join(root, "oracle"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_uniforms_open_at_zero(cref):
    u = cref.uniforms(1234, 4096)
    assert u.min() > 0.0 and u.max() <= 1.0  # Log(z1) must stay finite (FFTMesh.cs:173)
    assert abs(float(u.mean()) - 0.5) < 0.01


# ---------------------------------------------------------------- known answers from the source
@pytest.mark.parametrize("N", [16, 64])
def test_phillips_dc_and_symmetry(cref, N):
    p = cref.params(N)
    assert cref.phillips(p, N // 2, N // 2) == 0.0  # FFTMesh.cs:153-154
    rng = np.random.default_rng(0)
    for n, m in rng.integers(0, N, (50, 2)):
        assert cref.phillips(p, int(n), int(m)) == cref.phillips(p, N - int(n), N - int(m))  # P(k) == P(-k)
        assert cref.phillips(p, int(n), int(m)) >= 0.0


def test_phillips_suppresses_waves_across_the_wind(cref):
    p = cref.params(64, wind=(5.0, 0.0))
    # k perpendicular to the wind: (k.w)^2 == 0  (FFTMesh.cs:158-159)
    assert cref.phillips(p, 32, 40) == 0.0
    assert cref.phillips(p, 40, 32) > 0.0


def test_dispersion_is_quantised(cref, r64):
    N = 64
    p = cref.params(N)
    om = cref.dispersion(p)
    w0 = np.float32(2) * r64.PI / np.float32(p.length)
    q = om / w0
    assert np.abs(q - np.round(q)).max() < 1e-4  # multiples of w0 = 2 pi / L (FFTMesh.cs:146)
    assert np.array_equal(om.view(np.uint32), r64.omega_f32(N, p.length).view(np.uint32))  # numpy form is bit-exact
    assert om[N // 2, N // 2] == 0.0


def test_htilde_at_t0_is_h0_plus_h0conj(cref):
    p = cref.params(32)
    _, h0, hc = cref.generate_mesh(p, seed=5)
    H = cref.htilde(p, h0, hc, 0.0).reshape(-1, 2)
    assert np.array_equal(H, h0 + hc)  # cos 0 = 1, sin 0 = 0 exactly (FFTMesh.cs:183-188)


def test_whitecap_edges_and_range(cref):
    p = cref.params(16)
    v, h0, hc = cref.generate_mesh(p, seed=2)
    r = cref.evaluate_waves(p, v, h0, hc, 0.8, threads=2)
    hds = r["hds"].reshape(16, 16, 2)
    jac = r["jacobian"].reshape(16, 16)
    # last row: dDdx = 0 -> J = 1 + dDdy.y ; last column: dDdy = 0 -> J = 1 + dDdx.x ; corner: J = 1
    assert jac[15, 15] == 1.0
    assert np.allclose(jac[15, :15], 1 + 0.5 * (hds[15, :15, 1] - hds[15, 1:, 1]), atol=1e-6)
    assert np.allclose(jac[:15, 15], 1 + 0.5 * (hds[:15, 15, 0] - hds[1:, 15, 0]), atol=1e-6)
    w = r["whitecap"]
    assert w.min() >= 0.0 and w.max() <= 1.0
    assert np.array_equal(r["colors"][:, 0], r["colors"][:, 3])


def test_omp_split_matches_single_thread(cref):
    p = cref.params(16)
    v, h0, hc = cref.generate_mesh(p, seed=3)
    a = cref.evaluate_waves(p, v, h0, hc, 1.1, threads=1)
    b = cref.evaluate_waves(p, v, h0, hc, 1.1, threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k])
    vm, nr, hd = cref.evaluate_vertices(p, v, h0, hc, 1.1, 37, 53, threads=1)
    assert np.array_equal(vm, a["vertMeow"][37:53]) and np.array_equal(hd, a["hds"][37:53])


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["fftmesh_n32.npz", "fftmesh_n32_wind.npz", "fftmesh_n64.npz"])
def test_literal_oracle_reproduces_golden(cref, name):
    g = golden(name)
    N = int(g["N"])
    p = cref.params(N, 1.0, float(N), float(g["choppiness"]), float(g["amplitude"]), tuple(g["wind"]))
    v, h0, hc = cref.generate_mesh(p, seed=int(g["seed"]))
    assert np.array_equal(h0, g["h0"]) and np.array_equal(hc, g["h0conj"]) and np.array_equal(v, g["vertices"])
    assert np.array_equal(cref.dispersion(p), g["omega"])
    ts = g["ts"] if N <= 32 else g["ts"][:1]
    for k, t in enumerate(ts):
        r = cref.evaluate_waves(p, v, h0, hc, float(t), threads=cref.max_threads())
        for key in ("vertMeow", "normals", "hds", "jacobian", "whitecap"):
            assert np.array_equal(r[key], g[f"{key}_{k}"]), (name, key, k)


# ---------------------------------------------------------------- transform form == literal loop
@pytest.mark.parametrize("N,t", [(8, 0.0), (16, 1.7), (32, 60.0)])
def test_fft64_form_matches_literal(cref, r64, N, t):
    p = cref.params(N)
    v, h0, hc = cref.generate_mesh(p, seed=1234)
    lit = cref.evaluate_waves(p, v, h0, hc, t, threads=cref.max_threads())
    f = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, t)
    # the literal path accumulates N^2 fp32 terms with fp32 phases: ~1e-6 relative
    for a, b in (("height", "height"), ("hds", "hds"), ("normals", "normals"), ("vertMeow", "vertMeow"),
                 ("jacobian", "jacobian"), ("whitecap", "whitecap")):
        assert rel_l2(lit[a], f[b]) < 1e-5, (a, rel_l2(lit[a], f[b]))


def test_fft64_form_matches_golden_n64(r64):
    g = golden("fftmesh_n64.npz")
    f = r64.evaluate_waves(g["h0"], g["h0conj"], 64, 64.0, 1.0, 1.0, float(g["ts"][0]))
    assert rel_l2(g["hds_0"], f["hds"]) < 1e-5 and rel_l2(g["normals_0"], f["normals"]) < 1e-5
    assert max_abs(g["whitecap_0"], f["whitecap"]) < 1e-4


def test_ifft_identity_against_fp64_direct_sum(r64):
    """S = sigma N^2 ifft2(G r r) against the literal double loop in fp64."""
    rng = np.random.default_rng(0)
    for N in (8, 16):
        G = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        a = r64.direct_transform(G)
        b = r64.direct_sum_fp64(G, N, float(N) * 0.75, 0.75)
        assert np.abs(a - b).max() < 1e-11 * N * N


def test_identity_needs_periodic_sampling(r64):
    """FFT Mesh scene values (N=12, L=12.39) are NOT periodic: the identity must fail there, which is
    why the engine rejects length != resolution * unit_width instead of silently differing."""
    rng = np.random.default_rng(1)
    N = 12
    G = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    assert np.abs(r64.direct_transform(G) - r64.direct_sum_fp64(G, N, 12.39, 1.0)).max() > 1e-2


# ---------------------------------------------------------------- Stockham.shader restatement
@pytest.mark.parametrize("N", [8, 64, 256])
def test_stockham_stage_chain_is_forward_dft(r64, N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal((3, N, N)) + 1j * rng.standard_normal((3, N, N))
    y = r64.stockham_fft2d(x)
    assert np.abs(y - np.fft.fft2(x)).max() < 1e-9 * N


def test_stockham_pass_count_matches_ocean_renderer(r64):
    """OceanRenderer.cs:231: iterations = ceil(log2(resolution * 8)) * 2 = 20 blits per transform at
    R = 1024 (10 horizontal + 10 vertical); the restatement runs exactly that many stages."""
    R = 128 * 8
    assert int(np.ceil(np.log2(R))) * 2 == 20
    calls = []
    orig = r64.stockham_stage
    r64.stockham_stage = lambda x, sub, axis=-1: (calls.append((sub, axis)), orig(x, sub, axis))[1]
    try:
        r64.stockham_fft2d(np.zeros((16, 16), complex))
    finally:
        r64.stockham_stage = orig
    assert [c[0] for c in calls] == [2, 4, 8, 16] * 2 and [c[1] for c in calls] == [-1] * 4 + [-2] * 4


# ---------------------------------------------------------------- closed forms derived from the C# source, not from the oracle
def _plane_wave_closed_form(N, L, uw, chop, modes, t):
    """FFTMesh.cs:192-276 evaluated by hand for a spectrum with a handful of non-zero entries, in float64.

    modes = [(n, m, h0 (complex), h0conj (complex))].  For one entry, htilde = h0 e^{i w t} + h0conj e^{-i w t} (:178-190) and the
    double loop of Displacement (:199-217) has a single term: with phi = kx x + kz z,
        c      = htilde e^{i phi}
        height = Re c                                   (:210, :219)
        n     += (-kx Im c, 0, -kz Im c)  ->  normal = normalize((kx Im c, 1, kz Im c))   (:211, :218)
        d     += (kx / |k| Im c, -kz / |k| Im c)        (:214)
    so hds = d, vertMeow = (x - d.x chop, height, z - d.y chop) (:243-247); then the forward differences of :253-273."""
    half = N // 2
    off = uw / 2.0 if N % 2 == 0 else 0.0
    ax = (np.arange(N) - half) * uw + off                       # :107-112
    X, Z = np.meshgrid(ax, ax, indexing="ij")                   # index = i * N + j, i <-> x, j <-> z
    height = np.zeros((N, N)); dx = np.zeros((N, N)); dz = np.zeros((N, N)); sx = np.zeros((N, N)); sz = np.zeros((N, N))
    w0 = 2.0 * np.pi / L
    for n, m, h0, hc in modes:
        kx = 2.0 * np.pi * (n - N / 2.0) / L                    # :201, :204
        kz = 2.0 * np.pi * (m - N / 2.0) / L
        kl = np.hypot(kx, kz)
        om = np.floor(np.sqrt(9.81 * kl) / w0) * w0             # :141-147
        c = (h0 * np.exp(1j * om * t) + hc * np.exp(-1j * om * t)) * np.exp(1j * (kx * X + kz * Z))
        height += c.real
        sx += kx * c.imag; sz += kz * c.imag                   # nor = normalize(up - n), n = (-kx Im c, 0, -kz Im c)
        if kl >= 1e-4:                                          # :212
            dx += kx / kl * c.imag; dz += -kz / kl * c.imag
    inv = 1.0 / np.sqrt(sx * sx + 1.0 + sz * sz)
    normal = np.stack([sx * inv, inv, sz * inv], -1)
    vert = np.stack([X - dx * chop, height, Z - dz * chop], -1)
    ddx = np.zeros((N, N, 2)); ddy = np.zeros((N, N, 2))
    hds = np.stack([dx, dz], -1)
    ddx[:-1] = 0.5 * (hds[:-1] - hds[1:])                       # :260-263 (index + resolution = next i)
    ddy[:, :-1] = 0.5 * (hds[:, :-1] - hds[:, 1:])              # :264-267 (index + 1 = next j)
    jac = (1 + ddx[..., 0]) * (1 + ddy[..., 1]) - ddx[..., 1] * ddy[..., 0]
    noise = 0.3 * np.hypot(normal[..., 0], normal[..., 2])      # :269
    turb = np.maximum(1.0 - jac + noise, 0.0)
    s = np.clip(turb, 0.0, 1.0)
    white = s * s * (3.0 - 2.0 * s)                             # Mathf.SmoothStep(0, 1, turb), :273
    return {"vertMeow": vert.reshape(-1, 3), "normals": normal.reshape(-1, 3), "hds": hds.reshape(-1, 2), "jacobian": jac.reshape(-1),
            "whitecap": white.reshape(-1)}


@pytest.mark.parametrize("t", [0.0, 1.7, 60.0])
@pytest.mark.parametrize("modes", [
    [(20, 11, 0.35 - 0.2j, 0.0)],                                           # one travelling wave
    [(20, 11, 0.0, 0.25 + 0.4j)],                                           # only the conjugate partner: e^{-i w t}
    [(9, 25, 0.3 + 0.1j, -0.2 + 0.15j), (16, 16, 0.5 + 0.0j, 0.1 - 0.1j), (31, 0, -0.15 + 0.2j, 0.05j)],   # a sum, incl. k = 0
], ids=["h0", "h0conj", "three-modes-with-dc"])
def test_literal_oracle_against_closed_form_plane_waves(cref, modes, t):
    """The oracle's EvaluateWaves against a derivation made by hand from FFTMesh.cs for spectra with 1-3 non-zero entries: every
    output (displaced vertex, normal, hds, Jacobian, whitecap) in closed form.  Catches a mis-restated sign, index order (i <-> x),
    dispersion, conjugate term or forward-difference direction, which the oracle-vs-oracle fixtures cannot."""
    N = 32
    p = cref.params(N, choppiness=0.8)
    v, _, _ = cref.generate_mesh(p, seed=1)
    h0 = np.zeros((N * N, 2), np.float32); hc = np.zeros((N * N, 2), np.float32)
    for n, m, a, b in modes:
        h0[n * N + m] = (np.real(a), np.imag(a)); hc[n * N + m] = (np.real(b), np.imag(b))
    got = cref.evaluate_waves(p, v, h0, hc, t, threads=2)
    want = _plane_wave_closed_form(N, float(p.length), float(p.unit_width), 0.8, modes, t)
    # fp32 phases: |k.x| reaches ~100 rad and w t ~200 rad at t = 60 (one ulp of the angle is ~1e-5), amplitudes are O(1)
    tol = 2e-4 if t > 10 else 5e-5
    for key in ("vertMeow", "normals", "hds", "jacobian"):
        assert max_abs(got[key], want[key]) < tol, (key, max_abs(got[key], want[key]))
    assert max_abs(got["colors"][:, 0], want["whitecap"]) < 4 * tol


@pytest.mark.parametrize("N,L", [(12, 12.39), (13, 13.0)], ids=["fft-mesh-demo-scene", "odd-grid"])
def test_literal_oracle_against_closed_form_off_the_transform_grid(cref, N, L):
    """The same closed form on the grids the direct-sum path serves (the FFT Mesh scene's own resolution 12 / length 12.39,
    FFT Mesh.unity:147,150, and an odd grid, whose rest positions carry no half-cell offset, FFTMesh.cs:112): the sum is literal,
    so no periodicity is needed for the closed form to hold."""
    modes = [(2, 9, 0.3 + 0.1j, -0.2 + 0.15j), (N // 2, N // 2, 0.4 + 0.0j, 0.0), (N - 1, 0, -0.15 + 0.2j, 0.05j)]
    p = cref.params(N, length=L, choppiness=1.0)
    v, _, _ = cref.generate_mesh(p, seed=1)
    h0 = np.zeros((N * N, 2), np.float32); hc = np.zeros((N * N, 2), np.float32)
    for n, m, a, b in modes:
        h0[n * N + m] = (np.real(a), np.imag(a)); hc[n * N + m] = (np.real(b), np.imag(b))
    for t in (0.0, 1.7):
        got = cref.evaluate_waves(p, v, h0, hc, t, threads=1)
        want = _plane_wave_closed_form(N, L, 1.0, 1.0, modes, t)
        for key in ("vertMeow", "normals", "hds", "jacobian"):
            assert max_abs(got[key], want[key]) < 5e-5, (key, t, max_abs(got[key], want[key]))
        assert max_abs(got["colors"][:, 0], want["whitecap"]) < 2e-4
