"""Host-side mirror of the reference's OceanRenderer MonoBehaviour (Scripts/OceanRenderer.cs) -- the GPU-shader
convention the Ocean Demo scene runs -- over the mw_renderer_* C ABI (include/mistral_ocean.h).

`Renderer` is the thin handle wrapper (one C-ABI call per method); `OceanRenderer` has the MonoBehaviour's public
fields (OceanRenderer.cs:10-19), lifecycle (Awake / Update) and private method names (SetParams / GenerateMesh /
RenderInitial / GenerateTexture), with the RenderTextures exposed as numpy images [R, R, 4].  Nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import native
from .fft_mesh import Mesh
from .native import RendererOut, RendererParams, check
from .ocean import _addr

MAP_COMPONENTS = {"displacement": 4, "height": 4, "normal": 4, "white": 1, "white_rgba": 4, "jacobian": 1}


class Renderer:
    """One mw_renderer handle: `tiles` independent R x R oceans (R = 8 * resolution) on one GPU."""

    def __init__(self, resolution: int, length: float, choppiness: float = 1.5, amplitude: float = 1.0, wind=(1.0, 1.0),
                 mult: float = 2.0, unit_width: float = 1.0, seed1: float = 1.5122, seed2: float = 6.1152, device: int = 0,
                 tiles: int = 1, device_ptrs: bool = False, wrap_repeat: bool = False):
        self._lib = native.load()
        self._h = C.c_void_p()
        flags = (native.MW_DEVICE_PTRS if device_ptrs else 0) | (native.MW_WRAP_REPEAT if wrap_repeat else 0)
        self.params = RendererParams(int(resolution), float(unit_width), float(length), float(choppiness), float(amplitude),
                                     float(wind[0]), float(wind[1]), float(mult), float(seed1), float(seed2), int(device),
                                     int(tiles), flags, 0)
        check(self._lib.mw_renderer_create(C.byref(self.params), C.byref(self._h)))
        self.R = 8 * int(resolution)
        self.tiles = int(tiles)
        self.device_ptrs = device_ptrs

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mw_renderer_destroy(self._h)
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

    def render_initial(self) -> None:
        check(self._lib.mw_renderer_render_initial(self._h))

    def set_initial(self, rgba) -> None:
        check(self._lib.mw_renderer_set_initial(self._h, _addr(rgba)))

    def get_initial(self) -> np.ndarray:
        img = np.empty((self.tiles, self.R, self.R, 4), np.float32)
        check(self._lib.mw_renderer_get_initial(self._h, _addr(img)))
        return img

    def set_phase(self, phase) -> None:
        check(self._lib.mw_renderer_set_phase(self._h, _addr(phase)))

    def get_phase(self) -> np.ndarray:
        img = np.empty((self.tiles, self.R, self.R), np.float32)
        check(self._lib.mw_renderer_get_phase(self._h, _addr(img)))
        return img

    def set_params(self, length: float, choppiness: float, amplitude: float, wind) -> None:
        check(self._lib.mw_renderer_set_params(self._h, float(length), float(choppiness), float(amplitude),
                                               float(wind[0]), float(wind[1])))

    def alloc_maps(self, names=("displacement", "height", "normal", "white")) -> dict:
        return {k: np.empty((self.tiles, self.R, self.R, MAP_COMPONENTS[k]), np.float32) for k in names}

    def generate_texture(self, delta_time: float, bufs: dict | None = None,
                         names=("displacement", "height", "normal", "white")) -> dict:
        if bufs is None:
            bufs = self.alloc_maps(names)
        blk = RendererOut()
        for k in MAP_COMPONENTS:
            setattr(blk, k, _addr(bufs.get(k)) or None)
        check(self._lib.mw_renderer_generate_texture(self._h, float(delta_time), C.byref(blk)))
        return bufs

    def sync(self) -> None:
        check(self._lib.mw_renderer_sync(self._h))


def generate_mesh(resolution: int, unit_width: float, device: int = 0) -> Mesh:
    """mw_mesh_generate: GenerateMesh (OceanRenderer.cs:172-207 / FFTMesh.cs:101-139) -> vertices, normals, uv, indices."""
    n = int(resolution)
    m = Mesh(vertices=np.empty((n * n, 3), np.float32), normals=np.empty((n * n, 3), np.float32),
             uv=np.empty((n * n, 2), np.float32), indices=np.empty(((n - 1) * (n - 1) * 6,), np.int32))
    check(native.load().mw_mesh_generate(device, n, float(unit_width), m.vertices.ctypes.data, m.normals.ctypes.data,
                                         m.uv.ctypes.data, m.indices.ctypes.data))
    return m


@dataclass
class OceanRenderer:
    # ---- Public Variables, OceanRenderer.cs:10-19 (same names, same defaults) ----
    mult: float = 2.0
    unitWidth: float = 1.0
    resolution: int = 256
    length: float = 256.0
    choppiness: float = 1.5
    amplitude: float = 1.0
    wind: tuple = (0.0, 0.0)
    # ---- not in the reference: which GPU, the two Random.value * 10 draws of :147-148, the wrap mode ----
    device: int = 0
    randomSeed1: float = 1.5122   # InitialSpectrum.shader:6 default
    randomSeed2: float = 6.1152   # :7
    wrapRepeat: bool = False
    mesh: Mesh = field(default_factory=Mesh)

    def __post_init__(self):
        self._r: Renderer | None = None
        self.displacementTexture = self.heightTexture = self.normalTexture = self.whiteTexture = None

    # OceanRenderer.cs:76-89
    def Awake(self, initialTexture=None) -> None:
        self.SetParams()
        self.GenerateMesh()
        self.RenderInitial(initialTexture)

    # OceanRenderer.cs:91-110
    def Update(self, deltaTime: float) -> None:
        self.GenerateTexture(deltaTime)
        if (self._old[0] != self.length or self._old[2] != tuple(self.wind) or self._old[1] != self.amplitude):
            self._r.set_params(self.length, self.choppiness, self.amplitude, self.wind)  # :98-109 (re-renders the initial spectrum)
        elif self._old[3] != self.choppiness:
            self._r.set_params(self.length, self.choppiness, self.amplitude, self.wind)  # :96
        self._old = (self.length, self.amplitude, tuple(self.wind), self.choppiness)

    # OceanRenderer.cs:116-170
    def SetParams(self) -> None:
        if self._r is not None:
            self._r.close()
        self._r = Renderer(self.resolution, self.length, self.choppiness, self.amplitude, self.wind, self.mult,
                           self.unitWidth, self.randomSeed1, self.randomSeed2, self.device, wrap_repeat=self.wrapRepeat)
        self._old = (self.length, self.amplitude, tuple(self.wind), self.choppiness)

    # OceanRenderer.cs:172-207
    def GenerateMesh(self) -> None:
        self.mesh = generate_mesh(self.resolution, self.unitWidth, self.device)

    # OceanRenderer.cs:209-214
    def RenderInitial(self, initialTexture=None) -> None:
        if initialTexture is None:
            self._r.render_initial()
        else:
            self._r.set_initial(np.ascontiguousarray(initialTexture, np.float32))

    @property
    def initialTexture(self) -> np.ndarray:
        return self._r.get_initial()[0]

    # OceanRenderer.cs:216-316
    def GenerateTexture(self, deltaTime: float) -> None:
        m = self._r.generate_texture(deltaTime)
        self.displacementTexture, self.heightTexture = m["displacement"][0], m["height"][0]   # _Anim, _Height (:310, :313)
        self.normalTexture, self.whiteTexture = m["normal"][0], m["white"][0, ..., 0]         # _Bump, _White (:311-312)

    def close(self) -> None:
        if self._r is not None:
            self._r.close()
            self._r = None
