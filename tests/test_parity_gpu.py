"""GPU parity tests (run with -m gpu on a B200): the CUDA path, through the C ABI, against the CPU
oracle on the same seeded inputs; against the committed golden fixtures; and -- at BASELINE's full
sizes -- against the fp64 transform form of the oracle plus size-independent properties.

Tolerances (floating point path; stated per SURVEY.md section 8c):
  * omega (Dispersion): bit-exact.
  * h0 from the device init vs the oracle's init: relative L2 <= 1e-6.
  * fields vs the fp64 oracle: relative L2 <= 1e-5, max-abs <= 1e-5 * max|field| (whitecap: <= 1e-5 of
    the Jacobian's scale, it is a clamped smoothstep of a difference of large numbers; normals: max-abs
    <= 1e-4, a unit vector built from slopes of magnitude up to ~1e3 whose fp32 FFT error is absolute).
  * fields vs the literal fp32 O(N^4) oracle (N <= 64): relative L2 <= 2e-5, max-abs <= 1e-4 * max|field|
    -- the literal loop itself carries ~1e-6..1e-5 of fp32 phase/accumulation error (test_oracle.py).
"""
import numpy as np
import pytest

from conftest import golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu

PAIRS = (("height", "height"), ("disp", "hds"), ("normal", "normals"), ("vertices", "vertMeow"),
         ("jacobian", "jacobian"))
ALL = ("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors")


def _check_vs(out, ref, rel_tol, abs_frac, white_tol, normal_abs=None):
    for k, rk in PAIRS:
        r = rel_l2(out[k][0], ref[rk])
        assert r <= rel_tol, (k, r)
        lim = abs_frac * max(1.0, float(np.abs(ref[rk]).max()))
        if k == "normal" and normal_abs is not None:
            lim = normal_abs
        assert max_abs(out[k][0], ref[rk]) <= lim, (k, max_abs(out[k][0], ref[rk]), lim)
    assert max_abs(out["whitecap"][0], ref["whitecap"]) <= white_tol
    assert np.array_equal(out["colors"][0][:, 0], out["whitecap"][0][:, 0])
    assert np.array_equal(out["colors"][0][:, 0], out["colors"][0][:, 3])


# ------------------------------------------------------------------ a3-a6: init, dispersion
@pytest.mark.parametrize("N", [32, 64, 256, 1024])
def test_init_spectrum_and_dispersion(mw, cref, r64, N):
    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=1234)
    with mw.Ocean(N, seed=1234) as o:
        o.init_spectrum()
        g0, gc = o.get_h0()
        assert rel_l2(g0, h0) <= 1e-6 and rel_l2(gc, hc) <= 1e-6
        assert np.array_equal(o.dispersion().view(np.uint32), r64.omega_f32(N, p.length).view(np.uint32))
        assert np.array_equal(o.rest_vertices(), cref.generate_mesh(p, seed=0)[0])


def test_init_other_wind_amplitude(mw, cref):
    g = golden("fftmesh_n32_wind.npz")
    with mw.Ocean(32, amplitude=float(g["amplitude"]), wind=tuple(g["wind"]), seed=int(g["seed"])) as o:
        o.init_spectrum()
        g0, gc = o.get_h0()
        assert rel_l2(g0, g["h0"]) <= 1e-6 and rel_l2(gc, g["h0conj"]) <= 1e-6
        assert np.array_equal(o.dispersion(), g["omega"])


# ------------------------------------------------------------------ a7: htilde(t)
@pytest.mark.parametrize("t", [0.0, 1.7, 60.0, 3600.0])
def test_evolve_spectrum(mw, cref, t):
    N = 64
    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=7)
    with mw.Ocean(N) as o:
        o.set_h0(h0, hc)
        H = o.evolve_spectrum(t)[0]
        lit = cref.htilde(p, h0, hc, t)
        # same fp32 omega*t; sincosf vs (float)cos(double): <= 2 ulp of the trig factors
        assert max_abs(H, lit) <= 4e-7 * max(1.0, float(np.abs(lit).max()))
        if t == 0.0:
            assert np.array_equal(H.reshape(-1, 2), h0 + hc)


# ------------------------------------------------------------------ config 1: 64x64 vs the literal oracle
@pytest.mark.parametrize("t", [0.0, 1.7, 60.0])
def test_config1_64_vs_literal_oracle(mw, cref, r64, t):
    N = 64
    p = cref.params(N)  # L = 64, uw = 1, wind (5,3), A = 0.01, choppiness 1
    v, h0, hc = cref.generate_mesh(p, seed=1234)
    lit = cref.evaluate_waves(p, v, h0, hc, t, threads=cref.max_threads())
    f64 = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, t)
    with mw.Ocean(N, seed=1234) as o:
        o.set_h0(h0, hc)
        out = o.generate(t, names=ALL)
    _check_vs(out, lit, 2e-5, 1e-4, 1e-4)
    _check_vs(out, f64, 1e-5, 1e-5, 2e-5)


@pytest.mark.parametrize("name", ["fftmesh_n32.npz", "fftmesh_n64.npz", "fftmesh_n32_wind.npz"])
def test_golden_fixtures(mw, name):
    g = golden(name)
    N = int(g["N"])
    with mw.Ocean(N, choppiness=float(g["choppiness"]), amplitude=float(g["amplitude"]), wind=tuple(g["wind"])) as o:
        o.set_h0(g["h0"], g["h0conj"])
        for k, t in enumerate(g["ts"]):
            H = o.evolve_spectrum(float(t))[0]
            assert max_abs(H, g[f"htilde_{k}"]) <= 4e-7 * max(1.0, float(np.abs(g[f"htilde_{k}"]).max()))
            out = o.generate(float(t), names=ALL)
            ref = {key: g[f"{key}_{k}"] for key in ("vertMeow", "normals", "hds", "jacobian", "whitecap")}
            ref["height"] = ref["vertMeow"][:, 1]
            _check_vs(out, ref, 2e-5, 1e-4, 1e-4)


# ------------------------------------------------------------------ host mirror (reads like a test of FFTMesh)
def test_fftmesh_component_lifecycle(mw, cref):
    N = 32
    fm = mw.FFTMesh(choppiness=0.8, tDivision=2.0, resolution=N, unitWidth=1.0, length=32.0, wind=(5.0, 3.0),
                    amplitude=0.01, seed=42)
    fm.Awake()
    p = cref.params(N, choppiness=0.8)
    v, h0, hc = cref.generate_mesh(p, seed=42)
    assert np.array_equal(fm.vertices, v)
    fm.Awake(verttilde=h0, vertConj=hc)           # host-supplied spectrum (keeps UnityEngine.Random on the C# side)
    assert np.array_equal(fm.verttilde, h0) and np.array_equal(fm.vertConj, hc)
    fm.Update(0.5)
    fm.Update(0.25)                               # timer = (0.5 + 0.25) / tDivision
    assert abs(fm.timer - 0.375) < 1e-7
    lit = cref.evaluate_waves(p, v, h0, hc, 0.375, threads=cref.max_threads())
    assert rel_l2(fm.mesh.vertices, lit["vertMeow"]) <= 2e-5
    assert rel_l2(fm.mesh.normals, lit["normals"]) <= 2e-5
    assert max_abs(fm.mesh.colors, lit["colors"]) <= 1e-4
    fm.generate = True                            # FFTMesh.cs:62-68: re-init and timer = 0
    fm.Update(0.1)
    assert abs(fm.timer - 0.05) < 1e-7
    fm.close()


# ------------------------------------------------------------------ configs 2, 3 and the 2048 tile: vs fp64 form
@pytest.mark.parametrize("N,t", [(128, 1.7), (256, 1.7), (512, 0.4), (1024, 1.7), (2048, 1.7)])
def test_full_sizes_vs_fp64_oracle(mw, cref, r64, N, t):
    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=1234)
    ref = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, t)
    with mw.Ocean(N, seed=1234) as o:
        o.set_h0(h0, hc)
        out = o.generate(t, names=ALL)
    jscale = max(1.0, float(np.abs(ref["jacobian"]).max()))
    _check_vs(out, ref, 1e-5, 1e-5, 1e-5 * jscale, normal_abs=1e-4)


@pytest.mark.parametrize("N", [32, 64, 256, 1024, 2048])
def test_self_mirrored_rows_and_columns(mw, r64, N):
    """Rows / columns 0 and N/2 mirror onto themselves under k -> -k (kd[0] is the Nyquist value on both sides), so
    the Hermitian packing takes its general form there.  With a Phillips spectrum those modes carry ~no energy at
    large N and a mistake would hide below every tolerance: here ALL the energy sits on them (plus a few ordinary
    modes so that both code paths meet in one frame)."""
    rng = np.random.default_rng(N)
    mask = np.zeros((N, N), bool)
    mask[0, :] = mask[N // 2, :] = mask[:, 0] = mask[:, N // 2] = True
    mask[rng.integers(0, N, 24), rng.integers(0, N, 24)] = True
    h0 = (rng.standard_normal((N, N, 2)) * mask[..., None]).astype(np.float32).reshape(N * N, 2)
    hc = (rng.standard_normal((N, N, 2)) * mask[..., None]).astype(np.float32).reshape(N * N, 2)
    L = float(N)
    ref = r64.evaluate_waves(h0, hc, N, L, 1.0, 1.0, 0.9)
    with mw.Ocean(N) as o:
        o.set_h0(h0, hc)
        out = o.generate(0.9, names=("height", "disp", "normal", "jacobian"))
    for k, rk in (("height", "height"), ("disp", "hds"), ("normal", "normals"), ("jacobian", "jacobian")):
        assert rel_l2(out[k][0], ref[rk]) <= 1e-5, (k, rel_l2(out[k][0], ref[rk]))


def test_config2_256_outputs_subset(mw, cref, r64):
    """Config 2 asks for height + displacement + normal only (no whitecap): other pointers NULL."""
    N = 256
    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=1234)
    ref = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, 2.5)
    with mw.Ocean(N) as o:
        o.set_h0(h0, hc)
        out = o.generate(2.5, names=("height", "disp", "normal"))
        only_h = o.generate(2.5, names=("height",))
        only_w = o.generate(2.5, names=("whitecap",))
    assert rel_l2(out["height"], ref["height"]) <= 1e-5 and rel_l2(out["disp"], ref["hds"]) <= 1e-5
    assert rel_l2(out["normal"], ref["normals"]) <= 1e-5
    assert np.array_equal(only_h["height"], out["height"])
    assert max_abs(only_w["whitecap"], ref["whitecap"]) <= 1e-5 * float(np.abs(ref["jacobian"]).max())


# ------------------------------------------------------------------ size-independent properties at full size
@pytest.mark.parametrize("N", [1024, 2048])
def test_properties_full_size(mw, N):
    names = ("height", "disp", "normal", "whitecap", "jacobian")
    with mw.Ocean(N, seed=99) as o:
        o.init_spectrum()
        h0, hc = o.get_h0()
        a = o.generate(0.7, names=names)
        # (1) idempotence / determinism
        b = o.generate(0.7, names=names)
        for k in names:
            assert np.array_equal(a[k], b[k]), k
        # (2) checkpoint round trip: get_h0 -> set_h0 is the identity on the state
        o.set_h0(h0, hc)
        c = o.generate(0.7, names=names)
        for k in names:
            assert np.array_equal(a[k], c[k]), k
        # (3) linearity of the synthesis in the spectrum: scaling h0 scales height and hds
        o.set_h0(2.0 * h0, 2.0 * hc)
        d = o.generate(0.7, names=names)
        assert rel_l2(d["height"], 2.0 * a["height"]) <= 1e-6 and rel_l2(d["disp"], 2.0 * a["disp"]) <= 1e-6
        # (4) the DC bin never contributes to hds (FFTMesh.cs:213-214) and Phillips(DC) = 0: mean height == 0
        assert abs(float(a["height"].astype(np.float64).mean())) <= 1e-3 * float(np.abs(a["height"]).max())
        # (5) unit normals, whitecap in [0,1], Jacobian edge rule (FFTMesh.cs:260-268)
        nrm = a["normal"][0].astype(np.float64)
        assert np.abs(np.linalg.norm(nrm, axis=1) - 1.0).max() <= 1e-5 and nrm[:, 1].min() > 0.0
        w = a["whitecap"]
        assert w.min() >= 0.0 and w.max() <= 1.0 + 2.4e-7  # -2t^3 + 3t^2 in fp32 can round 1 ulp above 1
        hds = a["disp"][0].reshape(N, N, 2).astype(np.float64)
        jac = a["jacobian"][0].reshape(N, N).astype(np.float64)
        assert jac[N - 1, N - 1] == 1.0
        assert np.allclose(jac[N - 1, : N - 1], 1 + 0.5 * (hds[N - 1, : N - 1, 1] - hds[N - 1, 1:, 1]), rtol=1e-5, atol=1e-3)
        assert np.allclose(jac[: N - 1, N - 1], 1 + 0.5 * (hds[: N - 1, N - 1, 0] - hds[1:, N - 1, 0]), rtol=1e-5, atol=1e-3)
        # (6) Parseval: sum |height|^2 over the grid == N^2/... of the packed spectrum energy is covered by the
        #     fp64 comparison; here: time reversal symmetry of the dispersion, h(k,-t) from swapped h0/h0conj
        o.set_h0(hc, h0)
        e = o.generate(-0.7, names=("height",))
        assert rel_l2(e["height"], a["height"]) <= 1e-6


def test_tiles_are_independent_and_match_single_handles(mw):
    N, T = 256, 5
    with mw.Ocean(N, seed=300, tiles=T) as o:
        o.init_spectrum()
        batch = o.generate(1.3)
    for k in (0, 3, 4):
        with mw.Ocean(N, seed=300 + k) as s:
            s.init_spectrum()
            one = s.generate(1.3)
        for name in ("height", "disp", "normal", "whitecap"):
            assert np.array_equal(batch[name][k], one[name][0]), (k, name)


def test_tile_group_pipeline_matches_single_handles(mw):
    """tiles > group size: the frame is issued group by group on two streams with a double-buffered
    intermediate.  Odd group count, repeated frames, every tile must equal its own handle."""
    N, T = 1024, 3
    with mw.Ocean(N, seed=500, tiles=T) as o:
        o.init_spectrum()
        first = o.generate(0.25)
        again = o.generate(0.25)
        other = o.generate(2.0)
    for name in first:
        assert np.array_equal(first[name], again[name]), name
        assert not np.array_equal(first[name], other[name]), name
    for k in range(T):
        with mw.Ocean(N, seed=500 + k) as s:
            s.init_spectrum()
            one = s.generate(0.25)
        for name in ("height", "disp", "normal", "whitecap"):
            assert np.array_equal(first[name][k], one[name][0]), (k, name)


def test_update_timer_semantics_through_the_abi(mw):
    """mw_ocean_update == FFTMesh.Update: timer += dt / tDivision; EvaluateWaves(timer)  (FFTMesh.cs:70-72)."""
    with mw.Ocean(64, seed=3, t_division=4.0) as o:
        o.init_spectrum()
        bufs = o.alloc_outputs()
        o.update(1.0, bufs)
        o.update(0.5, bufs)
        assert abs(o.timer - 0.375) < 1e-7
        ref = o.generate(0.375)
        for k in bufs:
            assert np.array_equal(bufs[k], ref[k]), k
        o.reset_timer()
        assert o.timer == 0.0


def test_smallest_and_jacobian_only(mw, cref, r64):
    N = 32
    p = cref.params(N)
    _, h0, hc = cref.generate_mesh(p, seed=8)
    ref = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, 0.9)
    with mw.Ocean(N) as o:
        o.set_h0(h0, hc)
        out = o.generate(0.9, names=("jacobian",))
    assert max_abs(out["jacobian"], ref["jacobian"]) <= 1e-5 * max(1.0, float(np.abs(ref["jacobian"]).max()))


def test_device_pointer_mode_matches_host_mode(mw):
    import torch
    N = 512
    with mw.Ocean(N, seed=11) as h:
        h.init_spectrum()
        host = h.generate(0.9)
    st = torch.cuda.Stream()
    with mw.Ocean(N, seed=11, device_ptrs=True) as d:
        d.set_stream(st.cuda_stream)
        d.init_spectrum()
        bufs = {"height": torch.empty(N * N, device="cuda"), "disp": torch.empty(2 * N * N, device="cuda"),
                "normal": torch.empty(3 * N * N, device="cuda"), "whitecap": torch.empty(N * N, device="cuda")}
        d.generate(0.9, bufs)
        d.sync()
        for k in bufs:
            assert np.array_equal(bufs[k].cpu().numpy().reshape(-1), host[k].reshape(-1)), k


def test_host_async_mode_matches_blocking_calls(mw):
    """MW_HOST_ASYNC: calls return at once, outputs leave on the copy stream; after mw_ocean_sync the pinned host arrays
    hold exactly what the blocking calls produce, also when frames are queued back to back with a new h0 each."""
    import torch
    N, T = 256, 3
    with mw.Ocean(N, seed=5, tiles=T) as o:
        o.init_spectrum()
        h0, hc = o.get_h0()
        want = [o.generate(0.1 * k) for k in range(3)]
        o.set_h0(2.0 * h0, hc)
        want.append(o.generate(0.3))
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float32).pin_memory()  # noqa: E731
    ph0, phc = pin(T, N * N, 2), pin(T, N * N, 2)
    ph0.copy_(torch.from_numpy(h0)); phc.copy_(torch.from_numpy(hc))
    ph0b = pin(T, N * N, 2); ph0b.copy_(torch.from_numpy(2.0 * h0))
    outs = [{k: pin(T, N * N, c) for k, c in (("height", 1), ("disp", 2), ("normal", 3), ("whitecap", 1))} for _ in range(4)]
    with mw.Ocean(N, seed=5, tiles=T, host_async=True) as a:
        a.set_h0(ph0, phc)
        for k in range(3):
            a.generate(0.1 * k, outs[k])
        a.set_h0(ph0b, phc)
        a.generate(0.3, outs[3])
        a.sync()
        for k in range(4):
            for name in outs[k]:
                assert np.array_equal(outs[k][name].numpy(), want[k][name]), (k, name)


def test_state_errors(mw):
    with mw.Ocean(64) as o:
        with pytest.raises(mw.native.MwError) as ei:
            o.generate(0.0)
        assert ei.value.code == mw.native.MW_E_STATE  # EvaluateWaves before GenerateMesh
        with pytest.raises(mw.native.MwError):
            o.kernel_times()  # not created with MW_PROFILE


# ------------------------------------------------------------------ a10: the Stockham transform itself
@pytest.mark.parametrize("N", [32, 64, 128, 256, 512, 1024, 2048])
def test_fft2d_matches_numpy_and_stockham_restatement(mw, r64, N):
    rng = np.random.default_rng(N)
    x = (rng.standard_normal((2, N, N)) + 1j * rng.standard_normal((2, N, N))).astype(np.complex64)
    fwd = mw.fft2d(x, -1)
    ref = np.fft.fft2(x.astype(np.complex128))
    assert rel_l2(fwd.view(np.float32), ref.astype(np.complex64).view(np.float32)) <= 2e-6
    inv = mw.fft2d(fwd, +1) / np.float32(N * N)
    assert rel_l2(inv.view(np.float32), x.view(np.float32)) <= 3e-6  # forward -> inverse round trip
    if N <= 256:
        sref = r64.stockham_fft2d(x[0].astype(np.complex128))  # Stockham.shader stage chain
        assert rel_l2(fwd[0].view(np.float32), sref.astype(np.complex64).view(np.float32)) <= 2e-6


def test_fft2d_impulse_and_linearity(mw):
    N = 256
    x = np.zeros((1, N, N), np.complex64)
    x[0, 3, 5] = 1.0
    y = mw.fft2d(x, -1)[0]
    n = np.arange(N)
    want = np.exp(-2j * np.pi * (3 * n[:, None] + 5 * n[None, :]) / N)
    assert np.abs(y - want).max() <= 2e-6


@pytest.mark.parametrize("N", [64, 256, 1024])
def test_graph_replay_equals_plain_launches(mw, N, monkeypatch):
    """Single-group frames are replayed from a CUDA graph from the third call with the same output pointers on (the `t`
    argument of k_phase_table is patched per frame): bit-identical to a handle with graphs off (MW_GRAPH=0)."""
    import torch
    times = [0.0, 0.5, 1.7, 3.0, 60.0, 0.25]
    names = ("height", "disp", "normal", "whitecap")
    comps = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1}

    def run(graph):
        monkeypatch.setenv("MW_GRAPH", "1" if graph else "0")
        o = mw.Ocean(N, seed=42, device_ptrs=True)
        o.init_spectrum()
        bufs = {k: torch.zeros(N * N * comps[k], device="cuda") for k in names}
        outs = []
        before = mw.native.launch_count()
        for t in times:
            o.generate(t, bufs)
            o.sync()
            outs.append({k: v.clone() for k, v in bufs.items()})
        launches = mw.native.launch_count() - before
        # a different set of pointers falls back to plain launches, then re-captures
        other = {k: torch.zeros_like(v) for k, v in bufs.items()}
        for t in times[:3]:
            o.generate(t, other)
        o.sync()
        for k in names:
            assert torch.equal(other[k], outs[2][k]), k
        o.close()
        return outs, launches

    plain, n_plain = run(False)
    graph, n_graph = run(True)
    # the launch counter counts the kernels a replayed graph runs; small single frames (N <= 256) evaluate the phases inside
    # pass 1 and are two kernels, not three (and are not graphed)
    assert n_plain == n_graph == (2 if N <= 256 else 3) * len(times)
    for a, b in zip(plain, graph):
        for k in names:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("N,tiles", [(32, 1), (64, 1), (256, 1), (256, 4), (128, 16)])
def test_inline_phase_frames_equal_table_frames(mw, N, tiles, monkeypatch):
    """Small single frames skip k_phase_table: pass 1 evaluates sincosf(fl(fl(q w0) t)) itself -- the expression the table
    holds -- so the outputs are bit-identical to the three-kernel frame (MW_INLINE_PHASE=0), at large t too."""
    outs = []
    for inline in ("1", "0"):
        monkeypatch.setenv("MW_INLINE_PHASE", inline)
        with mw.Ocean(N, seed=9, tiles=tiles) as o:
            o.init_spectrum()
            before = mw.native.launch_count()
            res = [o.generate(t) for t in (0.0, 1.7, 3600.0)]
            per_frame = (mw.native.launch_count() - before) // 3
            assert per_frame == (2 if inline == "1" else 3)
            outs.append(res)
    for a, b in zip(*outs):
        for k in a:
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("N,tiles", [(256, 3), (512, 2), (1024, 2), (2048, 1)])
def test_halo_free_pass2_equals_halo_pass2(mw, N, tiles, monkeypatch):
    """From N = 512 up pass 2 runs without the halo line: a slab's east neighbour column is handed over by the CTA that owns it
    (csrc/mw_cols_seam.cuh) instead of being transformed a second time.  Same arithmetic on the same inputs: the outputs are
    bit-identical to the halo kernel's (MW_SEAM=0), frame after frame (the hand-over flags are cleared by pass 1), for every
    output set that has a whitecap or a Jacobian -- and mw_ocean_sync reports no timed-out hand-over."""
    res = {}
    for seam in ("0", "1"):
        monkeypatch.setenv("MW_SEAM", seam)
        with mw.Ocean(N, seed=21, tiles=tiles) as o:
            o.init_spectrum()
            frames = [o.generate(t, names=("height", "disp", "normal", "whitecap", "jacobian")) for t in (0.0, 1.7, 60.0)]
            frames.append(o.generate(2.5, names=("whitecap",)))
            frames.append(o.generate(2.5, names=("height", "disp", "normal")))
            o.sync()
            res[seam] = frames
    for a, b in zip(res["0"], res["1"]):
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(a[k], b[k]), (N, k)
