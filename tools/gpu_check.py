"""Developer script (runs on the GPU box): per-stage error report of the CUDA path vs the oracles."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mistral_water_b200 as mw
from oracle import cref, ref_fft64 as r64


def rel(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)), float(np.abs(a - b).max()), float(np.abs(b).max())


def check_fft(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((2, n, n)) + 1j * rng.standard_normal((2, n, n))).astype(np.complex64)
    for sign in (-1, 1):
        y = mw.fft2d(x, sign)
        ref = np.fft.fft2(x.astype(np.complex128)) if sign < 0 else np.fft.ifft2(x.astype(np.complex128)) * n * n
        print(f"fft2d N={n} sign={sign:+d}: relL2={rel(y.view(np.float32), ref.astype(np.complex64).view(np.float32))[0]:.2e}")


def check_ocean(N, t, literal=False, seed=1234):
    p = cref.params(N)
    v, h0, hc = cref.generate_mesh(p, seed=seed)
    with mw.Ocean(N, seed=seed) as o:
        o.init_spectrum()
        g0, gc = o.get_h0()
        om = o.dispersion()
        print(f"N={N} init: h0 rel={rel(g0, h0)[0]:.2e} h0c rel={rel(gc, hc)[0]:.2e} omega bit-exact={np.array_equal(om.view(np.uint32), r64.omega_f32(N, p.length).view(np.uint32))}")
        o.set_h0(h0, hc)
        H = o.evolve_spectrum(t)
        Href = r64.htilde(h0, hc, N, p.length, t)
        print(f"   evolve t={t}: rel={rel(H[0].reshape(N,N,2), np.stack([Href.real, Href.imag], -1))[0]:.2e}")
        out = o.generate(t, names=("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors"))
        ref = r64.evaluate_waves(h0, hc, N, p.length, p.unit_width, p.choppiness, t)
        for k, rk in (("height", "height"), ("disp", "hds"), ("normal", "normals"), ("whitecap", "whitecap"),
                      ("jacobian", "jacobian"), ("vertices", "vertMeow"), ("colors", "colors")):
            r = rel(out[k][0], ref[rk])
            print(f"   {k:9s} vs fft64  relL2={r[0]:.2e} maxabs={r[1]:.2e} (max|ref|={r[2]:.2e})")
        if literal:
            lit = cref.evaluate_waves(p, v, h0, hc, t, threads=cref.max_threads())
            for k, rk in (("height", "height"), ("disp", "hds"), ("normal", "normals"), ("whitecap", "whitecap"),
                          ("jacobian", "jacobian"), ("vertices", "vertMeow")):
                r = rel(out[k][0], lit[rk])
                print(f"   {k:9s} vs literal relL2={r[0]:.2e} maxabs={r[1]:.2e}")


def timing():
    import torch
    st = torch.cuda.Stream()
    for N, tiles in ((256, 1), (256, 64), (1024, 1), (1024, 16), (2048, 1), (2048, 4)):
        o = mw.Ocean(N, seed=1, tiles=tiles, device_ptrs=True, profile=True)
        o.set_stream(st.cuda_stream)
        o.init_spectrum()
        n2 = N * N * tiles
        bufs = {"height": torch.empty(n2, device="cuda"), "disp": torch.empty(n2 * 2, device="cuda"),
                "normal": torch.empty(n2 * 3, device="cuda"), "whitecap": torch.empty(n2, device="cuda")}
        with torch.cuda.stream(st):
            for i in range(5): o.generate(0.1 * i, bufs)
            torch.cuda.synchronize()
            o.kernel_times(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 50
            e0.record(st)
            for i in range(K): o.generate(0.016 * i, bufs)
            e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        kt, kn = o.kernel_times()
        print(f"N={N} tiles={tiles}: {ms*1e3:.1f} us/frame  {n2/ms/1e6:.2f} Gpts/s  eff GB/s(44B)={n2*44/ms/1e6:.0f}  rows={kt[0]/max(kn[0],1)*1e3:.1f}us cols={kt[1]/max(kn[1],1)*1e3:.1f}us")
        o.close()


if __name__ == "__main__":
    if "--timing" in sys.argv:
        timing(); sys.exit(0)
    for n in (32, 64, 128, 256, 512, 1024, 2048):
        check_fft(n)
    for N in (32, 64, 128, 256, 512, 1024, 2048):
        check_ocean(N, 1.7, literal=N <= 64)
    check_ocean(64, 60.0, literal=True)
    check_ocean(64, 0.0, literal=True)
    # gerstner
    g = mw.pond_wave_table_32()
    N = 1024
    pos = np.zeros((N * N, 3), np.float32)
    ax = (np.arange(N) - N // 2 + 0.5).astype(np.float32)
    pos[:, 0] = np.repeat(ax, N); pos[:, 2] = np.tile(ax, N)
    out = g.displace(pos, 1.7)
    ref = cref.gerstner_table(g.table(), pos, 1.7)
    print("gerstner32 1M: maxabs", np.abs(out - ref).max(), "max|offs|", np.abs(ref - pos).max())
    timing()

