"""ctypes loader for oracle/libmw_oracle.so (the literal C restatement) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmw_oracle.so")


class RefParams(C.Structure):
    _fields_ = [("resolution", C.c_int32), ("unit_width", C.c_float), ("length", C.c_float),
                ("choppiness", C.c_float), ("amplitude", C.c_float), ("wind_x", C.c_float), ("wind_y", C.c_float)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("ref_fftmesh.c", "ref_gerstner.c", "ref_philox.h", "Makefile")]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libmw_oracle.so"], check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        fp = C.POINTER(C.c_float)
        P = C.POINTER(RefParams)
        _lib.ref_dispersion.restype = C.c_float
        _lib.ref_dispersion.argtypes = [P, C.c_int, C.c_int]
        _lib.ref_phillips.restype = C.c_float
        _lib.ref_phillips.argtypes = [P, C.c_int, C.c_int]
        _lib.ref_uniforms.argtypes = [C.c_uint64, C.c_int64, fp]
        _lib.ref_generate_mesh.argtypes = [P, fp, C.c_uint64, fp, fp, fp]
        _lib.ref_htilde.argtypes = [P, fp, fp, C.c_float, fp]
        _lib.ref_evaluate_vertices.argtypes = [P, fp, fp, fp, C.c_float, C.c_int64, C.c_int64, C.c_int, fp, fp, fp]
        _lib.ref_whitecaps.argtypes = [P, fp, fp, fp, fp]
        _lib.ref_evaluate_waves.argtypes = [P, fp, fp, fp, C.c_float, C.c_int, fp, fp, fp, fp, fp]
        _lib.ref_max_threads.restype = C.c_int
        _lib.ref_gerstner4.argtypes = [fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp, fp, fp]
        _lib.ref_gerstner_level_one.argtypes = [fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        _lib.ref_gerstner_table.argtypes = [fp, C.c_int, fp, C.c_int64, C.c_float, fp, fp]
        _lib.ref_gerstner_table_normals.argtypes = [fp, C.c_int, fp, C.c_int64, C.c_float, C.c_int, C.c_float, fp]
        _lib.ref_wave.argtypes = [fp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def params(resolution, unit_width=1.0, length=None, choppiness=1.0, amplitude=0.01, wind=(5.0, 3.0)) -> RefParams:
    if length is None:
        length = resolution * unit_width
    return RefParams(int(resolution), float(unit_width), float(length), float(choppiness), float(amplitude),
                     float(wind[0]), float(wind[1]))


def dispersion(p: RefParams) -> np.ndarray:
    N = p.resolution
    out = np.empty((N, N), np.float32)
    L = lib()
    for n in range(N):
        for m in range(N):
            out[n, m] = L.ref_dispersion(C.byref(p), n, m)
    return out


def phillips(p: RefParams, n: int, m: int) -> float:
    return float(lib().ref_phillips(C.byref(p), n, m))


def uniforms(seed: int, count: int) -> np.ndarray:
    out = np.empty((count, 4), np.float32)
    L = lib()
    buf = (C.c_float * 4)()
    for i in range(count):
        L.ref_uniforms(seed, i, buf)
        out[i] = buf[:]
    return out


def generate_mesh(p: RefParams, seed: int = 0, uniforms_: np.ndarray | None = None):
    """FFTMesh.GenerateMesh -> (vertices [N*N,3], h0 [N*N,2], h0conj [N*N,2])."""
    n2 = p.resolution * p.resolution
    v = np.empty((n2, 3), np.float32)
    h0 = np.empty((n2, 2), np.float32)
    hc = np.empty((n2, 2), np.float32)
    u = None if uniforms_ is None else _f32(uniforms_)
    lib().ref_generate_mesh(C.byref(p), _p(u), C.c_uint64(seed), _p(v), _p(h0), _p(hc))
    return v, h0, hc


def htilde(p: RefParams, h0, h0conj, t: float) -> np.ndarray:
    N = p.resolution
    out = np.empty((N, N, 2), np.float32)
    lib().ref_htilde(C.byref(p), _p(_f32(h0)), _p(_f32(h0conj)), t, _p(out))
    return out


def evaluate_waves(p: RefParams, vertices, h0, h0conj, t: float, threads: int = 1) -> dict:
    """FFTMesh.EvaluateWaves(t), literal O(N^4)."""
    n2 = p.resolution * p.resolution
    vm = np.array(vertices, dtype=np.float32, copy=True).reshape(n2, 3)
    nr = np.zeros((n2, 3), np.float32)
    hds = np.zeros((n2, 2), np.float32)
    jac = np.zeros((n2,), np.float32)
    col = np.zeros((n2, 4), np.float32)
    lib().ref_evaluate_waves(C.byref(p), _p(_f32(vertices)), _p(_f32(h0)), _p(_f32(h0conj)), t, threads,
                             _p(vm), _p(nr), _p(hds), _p(jac), _p(col))
    return {"vertMeow": vm, "normals": nr, "hds": hds, "jacobian": jac, "colors": col,
            "height": vm[:, 1].copy(), "whitecap": col[:, 0].copy()}


def evaluate_vertices(p: RefParams, vertices, h0, h0conj, t: float, v_begin: int, v_end: int, threads: int = 1):
    """First half of EvaluateWaves for a vertex range (bounded CPU-baseline samples)."""
    n2 = p.resolution * p.resolution
    vm = np.zeros((n2, 3), np.float32)
    nr = np.zeros((n2, 3), np.float32)
    hds = np.zeros((n2, 2), np.float32)
    lib().ref_evaluate_vertices(C.byref(p), _p(_f32(vertices)), _p(_f32(h0)), _p(_f32(h0conj)), t,
                                v_begin, v_end, threads, _p(vm), _p(nr), _p(hds))
    return vm[v_begin:v_end], nr[v_begin:v_end], hds[v_begin:v_end]


def whitecaps(p: RefParams, hds, normals):
    n2 = p.resolution * p.resolution
    jac = np.zeros((n2,), np.float32)
    col = np.zeros((n2, 4), np.float32)
    lib().ref_whitecaps(C.byref(p), _p(_f32(hds)), _p(_f32(normals)), _p(jac), _p(col))
    return jac, col


def max_threads() -> int:
    return int(lib().ref_max_threads())


def gerstner4(pos, t, amplitude, frequency, steepness, speed, dirAB, dirCD):
    pos = _f32(pos)
    out = np.empty_like(pos)
    lib().ref_gerstner4(_p(pos), pos.shape[0], t, amplitude, frequency, steepness,
                        _p(_f32(speed)), _p(_f32(dirAB)), _p(_f32(dirCD)), _p(out))
    return out


def gerstner_level_one(pos, t, amplitude, frequency, steepness):
    pos = _f32(pos)
    out = np.empty_like(pos)
    lib().ref_gerstner_level_one(_p(pos), pos.shape[0], t, amplitude, frequency, steepness, _p(out))
    return out


def gerstner_table(waves, pos, t, want_normal=False):
    pos = _f32(pos)
    waves = _f32(waves).reshape(-1, 6)
    out = np.empty_like(pos)
    nrm = np.empty_like(pos) if want_normal else None
    lib().ref_gerstner_table(_p(waves), waves.shape[0], _p(pos), pos.shape[0], t, _p(out), _p(nrm))
    return (out, nrm) if want_normal else out


def gerstner_table_normals(waves, pos, t, mode="analytic", smoothing=1.0):
    """The normals the reference computes but does not ship: "analytic" (the displaced surface's own normal, what
    MistralWaterLib.cginc:122-124 was after) or "discarded" (Gerstner() :92-97 literally)."""
    pos = _f32(pos)
    waves = _f32(waves).reshape(-1, 6)
    nrm = np.empty_like(pos)
    lib().ref_gerstner_table_normals(_p(waves), waves.shape[0], _p(pos), pos.shape[0], t, {"analytic": 2, "discarded": 3}[mode],
                                     smoothing, _p(nrm))
    return nrm


def wave(pos, t, amplitude, frequency, speed, smoothing):
    """MistralWaterLib.cginc:127-152 Wave through Displacement :160-164 -> (displaced vertices, normals)."""
    pos = _f32(pos)
    out = np.empty_like(pos)
    nrm = np.empty_like(pos)
    lib().ref_wave(_p(pos), pos.shape[0], t, amplitude, frequency, speed, smoothing, _p(out), _p(nrm))
    return out, nrm
