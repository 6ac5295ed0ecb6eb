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
for fl, name in ((0, "full"), (1, "AB: no stores"), (16, "AB: no extraction"), (2, "AB: no FFT"), (4, "AB: no loads"), (8, "no C"), (8 + 16, "no C, no extraction"), (8 + 16 + 2, "no C, only loads"), (8+16+2+4, "nothing"),
                 (32, "rows: no evolve"), (64, "rows: no fft/store")):
    r, c = timeit(fl)
    print(f"flags={fl:4d} ({name:36s}): rows {r:6.0f} us   cols {c:6.0f} us   per 16-tile frame")
