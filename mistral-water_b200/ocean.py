"""Thin object wrapper over the mw_ocean_* C ABI (include/mistral_ocean.h).

`Ocean` works on numpy host arrays (the default, what the C# host does with pinned managed arrays)
or, with device_ptrs=True, on raw device addresses / torch CUDA tensors (what bench.py and the
multi-GPU tile path use).  Nothing here computes: every method is one C-ABI call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import OceanOut, OceanParams, check

OUT_COMPONENTS = {"height": 1, "disp": 2, "normal": 3, "whitecap": 1, "jacobian": 1, "vertices": 3, "colors": 4}


def _addr(buf) -> int:
    """Address of a numpy array (host) or a torch tensor (host or device)."""
    if buf is None:
        return 0
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.float32 or not buf.flags["C_CONTIGUOUS"]:
            raise TypeError("buffers must be C-contiguous float32")
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):  # torch.Tensor
        if not buf.is_contiguous() or str(buf.dtype) != "torch.float32":
            raise TypeError("tensors must be contiguous float32")
        return buf.data_ptr()
    raise TypeError(f"unsupported buffer type {type(buf)!r}")


class Ocean:
    """One mw_ocean handle: `tiles` independent N x N oceans on one GPU."""

    def __init__(self, resolution: int, unit_width: float = 1.0, length: float | None = None,
                 choppiness: float = 1.0, amplitude: float = 0.01, wind=(5.0, 3.0), t_division: float = 1.0,
                 seed: int = 0, device: int = 0, tiles: int = 1, device_ptrs: bool = False, profile: bool = False,
                 host_async: bool = False):
        self._lib = native.load()
        self._h = C.c_void_p()
        if length is None:
            length = float(np.float32(resolution) * np.float32(unit_width))
        flags = ((native.MW_DEVICE_PTRS if device_ptrs else 0) | (native.MW_PROFILE if profile else 0) |
                 (native.MW_HOST_ASYNC if host_async else 0))
        self.params = OceanParams(int(resolution), float(unit_width), float(length), float(choppiness),
                                  float(amplitude), float(wind[0]), float(wind[1]), float(t_division),
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, int(device), int(tiles), flags, 0)
        check(self._lib.mw_ocean_create(C.byref(self.params), C.byref(self._h)))
        self.N = int(resolution)
        self.tiles = int(tiles)
        self.device_ptrs = device_ptrs

    # -- lifetime -----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mw_ocean_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state --------------------------------------------------------------------------------
    @property
    def points(self) -> int:
        return self.tiles * self.N * self.N

    def init_spectrum(self) -> None:
        check(self._lib.mw_ocean_init_spectrum(self._h))

    def set_h0(self, h0, h0conj) -> None:
        check(self._lib.mw_ocean_set_h0(self._h, _addr(h0), _addr(h0conj)))

    def get_h0(self):
        if self.device_ptrs:
            raise RuntimeError("get_h0() allocates host arrays; use get_h0_into() with device buffers")
        h0 = np.empty((self.tiles, self.N * self.N, 2), np.float32)
        hc = np.empty_like(h0)
        check(self._lib.mw_ocean_get_h0(self._h, _addr(h0), _addr(hc)))
        return h0, hc

    def get_h0_into(self, h0, h0conj) -> None:
        check(self._lib.mw_ocean_get_h0(self._h, _addr(h0), _addr(h0conj)))

    def rest_vertices(self) -> np.ndarray:
        v = np.empty((self.N * self.N, 3), np.float32)
        check(self._lib.mw_ocean_get_rest_vertices(self._h, _addr(v)))
        return v

    def dispersion(self) -> np.ndarray:
        w = np.empty((self.N, self.N), np.float32)
        check(self._lib.mw_ocean_get_dispersion(self._h, _addr(w)))
        return w

    def evolve_spectrum(self, t: float, out=None):
        if out is None:
            if self.device_ptrs:
                raise RuntimeError("pass a device buffer")
            out = np.empty((self.tiles, self.N, self.N, 2), np.float32)
        check(self._lib.mw_ocean_evolve_spectrum(self._h, float(t), _addr(out)))
        return out

    # -- per frame ----------------------------------------------------------------------------
    def _out_block(self, bufs: dict) -> OceanOut:
        blk = OceanOut()
        for k in OUT_COMPONENTS:
            setattr(blk, k, _addr(bufs.get(k)) or None)
        return blk

    def alloc_outputs(self, names=("height", "disp", "normal", "whitecap")) -> dict:
        """Host output arrays shaped [tiles, N*N, components]."""
        return {k: np.empty((self.tiles, self.N * self.N, OUT_COMPONENTS[k]), np.float32) for k in names}

    def generate(self, t: float, bufs: dict | None = None, names=("height", "disp", "normal", "whitecap")) -> dict:
        """mw_ocean_generate: EvaluateWaves(t).  bufs maps output name -> buffer (None: allocate host arrays)."""
        if bufs is None:
            bufs = self.alloc_outputs(names)
        blk = self._out_block(bufs)
        check(self._lib.mw_ocean_generate(self._h, float(t), C.byref(blk)))
        return bufs

    def update(self, delta_time: float, bufs: dict) -> dict:
        blk = self._out_block(bufs)
        check(self._lib.mw_ocean_update(self._h, float(delta_time), C.byref(blk)))
        return bufs

    def reset_timer(self) -> None:
        check(self._lib.mw_ocean_reset_timer(self._h))

    @property
    def timer(self) -> float:
        return float(self._lib.mw_ocean_timer(self._h))

    def sync(self) -> None:
        check(self._lib.mw_ocean_sync(self._h))

    def set_stream(self, cuda_stream: int | None) -> None:
        check(self._lib.mw_ocean_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def kernel_times(self, reset: bool = True):
        ms = (C.c_float * native.MW_KERNEL_COUNT)()
        n = (C.c_int64 * native.MW_KERNEL_COUNT)()
        check(self._lib.mw_ocean_kernel_times(self._h, ms, n, int(reset)))
        return list(ms), list(n)


def fft2d(x: np.ndarray, sign: int = -1, device: int = 0) -> np.ndarray:
    """mw_fft2d on a complex64 array [..., N, N] (host).  sign=-1 == numpy.fft.fft2."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    n = x.shape[-1]
    if x.ndim < 2 or x.shape[-2] != n:
        raise ValueError("expected [..., N, N]")
    batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    out = np.empty_like(x)
    check(native.load().mw_fft2d(device, n, batch, sign, x.ctypes.data, out.ctypes.data))
    return out
