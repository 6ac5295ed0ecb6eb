"""Host-side mirror of the pond renderer's Gerstner displacement
(Shaders/MistralWaterLib.cginc Displacement :154-180 -> Gerstner :71-99 / GerstnerLevelOne :101-125).
Material property names are the shader's (MistralWaterBasic.shader property block)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import GerstnerParams, check
from .ocean import _addr

# Materials/Pond Water Mat.mat:90,104,122,134-136
POND_MATERIAL = dict(_Amplitude=10.0, _Frequency=2.58, _Steepness=0.99,
                     _WSpeed=(1.2, 0.71, 1.1, 0.73), _WDirectionAB=(0.3, 0.73, 0.85, 0.25),
                     _WDirectionCD=(-0.25, 1.11, 0.5, 0.5))


def _f4(v):
    return (C.c_float * 4)(*[float(x) for x in v])


class GerstnerWaves:
    """A W-wave table (W <= 64) plus the displace call."""

    def __init__(self, device: int = 0, device_ptrs: bool = False):
        self._lib = native.load()
        self.p = GerstnerParams()
        self.p.n_waves = 0
        self.p.device = device
        self.p.flags = native.MW_DEVICE_PTRS if device_ptrs else 0

    @classmethod
    def from_material(cls, _Amplitude, _Frequency, _Steepness, _WSpeed, _WDirectionAB, _WDirectionCD, **kw):
        """The 4-wave Gerstner of the `_DISPLACEMENTMODE_GERSTNER` keyword (MistralWaterLib.cginc:168-177)."""
        g = cls(**kw)
        check(g._lib.mw_gerstner_from_material(C.byref(g.p), _Amplitude, _Frequency, _Steepness,
                                               _f4(_WSpeed), _f4(_WDirectionAB), _f4(_WDirectionCD)))
        return g

    def append_level_one(self, amplitude, frequency, steepness):
        """GerstnerLevelOne's five table waves (MistralWaterLib.cginc:105-109)."""
        check(self._lib.mw_gerstner_append_level_one(C.byref(self.p), amplitude, frequency, steepness))
        return self

    def append(self, dir_x, dir_y, freq, rate, amp_xz, amp_y):
        if self.p.n_waves >= native.MW_GERSTNER_MAX_WAVES:
            raise ValueError("wave table full")
        w = self.p.waves[self.p.n_waves]
        w.dir_x, w.dir_y, w.freq, w.rate, w.amp_xz, w.amp_y = map(float, (dir_x, dir_y, freq, rate, amp_xz, amp_y))
        self.p.n_waves += 1
        return self

    @property
    def n_waves(self) -> int:
        return int(self.p.n_waves)

    def table(self) -> np.ndarray:
        return np.array([[w.dir_x, w.dir_y, w.freq, w.rate, w.amp_xz, w.amp_y]
                         for w in self.p.waves[: self.p.n_waves]], np.float32).reshape(-1, 6)

    def displace(self, pos, t: float, out=None, normals=None, stream: int = 0, normal_mode: str = "up", smoothing: float = 1.0):
        """v.vertex.xyz += offsets (MistralWaterLib.cginc:176).  pos/out: [n, 3] float32.
        normals (optional [n, 3] buffer) receives, by normal_mode: "up" = (0, 1, 0), what the shader ships (:98, :121);
        "analytic" = the displaced surface's own normal (what the commented lines :122-124 were after); "discarded" = the
        value Gerstner() computes at :92-97 and then overwrites (uses _Smoothing)."""
        n = int(pos.shape[0])
        if out is None:
            out = np.empty_like(pos)
        bits = {"up": 0, "analytic": native.MW_GERSTNER_NORMAL_ANALYTIC, "discarded": native.MW_GERSTNER_NORMAL_DISCARDED}[normal_mode]
        self.p.flags = (self.p.flags & ~(native.MW_GERSTNER_NORMAL_ANALYTIC | native.MW_GERSTNER_NORMAL_DISCARDED)) | bits
        self.p.smoothing = float(smoothing)
        check(self._lib.mw_gerstner_displace(C.byref(self.p), _addr(pos), _addr(out), _addr(normals) or None,
                                             n, float(t), C.c_void_p(stream)))
        return out


def pond_wave_table_32(seed: int = 7, **kw) -> GerstnerWaves:
    """BASELINE config 4's 32-wave table: waves 0-3 = the Pond material's Gerstner, 4-8 =
    GerstnerLevelOne's tables (same material scalars), 9-31 = seeded perturbations in the same ranges."""
    m = POND_MATERIAL
    g = GerstnerWaves.from_material(**m, **kw)
    amp = m["_Amplitude"] * 0.01
    g.append_level_one(amp, m["_Frequency"], m["_Steepness"])
    rng = np.random.default_rng(seed)
    for _ in range(32 - g.n_waves):
        d = rng.uniform(-1.2, 1.2, 2)
        fs = rng.uniform(0.2, 1.6)
        sp = rng.uniform(-3.7, 1.3)
        a = rng.uniform(0.4, 0.9)
        st = rng.uniform(0.4, 0.95)
        f = m["_Frequency"] * fs
        g.append(d[0], d[1], f, sp * f, m["_Steepness"] * amp * st * a / 4.0, amp * a / 4.0)
    return g


def wave_displace(pos, t, _Amplitude, _Frequency, _Speed, _Smoothing, want_normal=True, device=0):
    """The `_DISPLACEMENTMODE_WAVE` branch of Displacement (MistralWaterLib.cginc:160-164 -> Wave :127-152) on host arrays:
    returns (displaced vertices, normals)."""
    pos = np.ascontiguousarray(pos, np.float32)
    out = np.empty_like(pos)
    nrm = np.empty_like(pos) if want_normal else None
    p = native.WaveParams(float(_Amplitude), float(_Frequency), float(_Speed), float(_Smoothing), int(device), 0)
    check(native.load().mw_wave_displace(C.byref(p), _addr(pos), _addr(out), _addr(nrm), pos.shape[0], float(t), None))
    return out, nrm
