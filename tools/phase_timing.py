"""Developer script: per-phase clock64 stamps of the two frame kernels (N=1024, 16 tiles)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mistral_water_b200 as mw
N, tiles = 1024, 16
lib = mw.native.load()
lib.mw_debug_phase_buffers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
st = torch.cuda.Stream()
o = mw.Ocean(N, seed=1, tiles=tiles, device_ptrs=True, profile=True)
o.set_stream(st.cuda_stream); o.init_spectrum()
n2 = N * N * tiles
bufs = {"height": torch.empty(n2, device="cuda"), "disp": torch.empty(n2 * 2, device="cuda"),
        "normal": torch.empty(n2 * 3, device="cuda"), "whitecap": torch.empty(n2, device="cuda")}
rows_ctas, cols_ctas = (N // 2) * tiles, (N // 4 + N // 8) * tiles
dr = torch.zeros(rows_ctas * 8, dtype=torch.int64, device="cuda")
dc = torch.zeros(cols_ctas * 8, dtype=torch.int64, device="cuda")
lib.mw_debug_flags.argtypes = [C.c_void_p, C.c_int]
def timeit(flags, K=20):
    lib.mw_debug_flags(o._h, flags)
    lib.mw_debug_phase_buffers(o._h, None, None)
    prof = []
    with torch.cuda.stream(st):
        for i in range(3): o.generate(0.1 * i, bufs)
        torch.cuda.synchronize()
        o.kernel_times(reset=True)
        for i in range(K): o.generate(0.1 * i, bufs)
        ms, n = o.kernel_times()
    return ms[0] / n[0] * 1e3, ms[1] / n[1] * 1e3
for fl, name in ((16384, "AB: contiguous (wrong-place) stores"), (16384 + 8, "same, no C"),
                 (0, "full"), (1, "AB: no output stores"), (2, "AB: no FFT"), (4, "AB: no slab load"), (7, "AB: only extract math"), (7 + 16, "AB: nothing, C full"), (8, "no C"), (8 + 7 + 16, "cols: only twiddle load + barriers"), (256, "cols: return at once"),
                 (32, "rows: no evolve"), (64, "rows: no fft/store"), (32 + 64, "rows: only twiddles"), (512, "rows: return before evolve")):
    r, c = timeit(fl)
    print(f"flags={fl:4d} ({name:36s}): rows {r:6.0f} us   cols {c:6.0f} us   per 16-tile frame")
lib.mw_debug_flags(o._h, 0)
with torch.cuda.stream(st):
    lib.mw_debug_phase_buffers(o._h, dr.data_ptr(), dc.data_ptr())
    o.generate(0.5, bufs); torch.cuda.synchronize()
r = dr.cpu().numpy().reshape(-1, 8); c = dc.cpu().numpy().reshape(tiles, -1, 8)
d = np.diff(r[:, :4], axis=1)
print("rows kernel, cycles per CTA: evolve %.0f  sync %.0f  fft+store %.0f   total %.0f" % (*d.mean(0), (r[:, 3] - r[:, 0]).mean()))
ab = c[:, : N // 4].reshape(-1, 8)
d = np.diff(ab[:, :6], axis=1)
print("cols AB CTA: cp.async+wait %.0f  sync %.0f  fft %.0f  sync %.0f  extract %.0f   total %.0f" % (*d.mean(0), (ab[:, 5] - ab[:, 0]).mean()))
print("  p10/p50/p90 total:", np.percentile(ab[:, 5] - ab[:, 0], [10, 50, 90]))
