"""ctypes binding of libmistral_ocean.so -- the same exported symbols the C# side reaches through
[DllImport("mistral_ocean")] (bindings/MistralOceanNative.cs).  There is no CPU path: if the
library is missing this module raises, and every call that needs a GPU fails with the library's
own error text."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", f"libmistral_ocean{os.environ.get('MW_LIB_SUFFIX', '')}.so")  # suffix: developer experiment builds

MW_OK = 0
MW_E_INVALID_ARG = -1
MW_E_CUDA = -2
MW_E_OOM = -3
MW_E_STATE = -4
MW_E_NCCL = -5
MW_DEVICE_PTRS = 1 << 0
MW_PROFILE = 1 << 1
MW_WRAP_REPEAT = 1 << 2
MW_HOST_ASYNC = 1 << 3
MW_GERSTNER_NORMAL_ANALYTIC = 1 << 4
MW_GERSTNER_NORMAL_DISCARDED = 1 << 5
MW_KERNEL_COUNT = 3
MW_GERSTNER_MAX_WAVES = 64

# every symbol include/mistral_ocean.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "mw_version", "mw_last_error", "mw_ocean_create", "mw_ocean_destroy", "mw_ocean_init_spectrum",
    "mw_ocean_set_h0", "mw_ocean_get_h0", "mw_ocean_get_rest_vertices", "mw_ocean_get_dispersion",
    "mw_ocean_evolve_spectrum", "mw_ocean_generate", "mw_ocean_update", "mw_ocean_reset_timer",
    "mw_ocean_timer", "mw_ocean_sync", "mw_ocean_set_stream", "mw_ocean_kernel_times",
    "mw_kernel_launch_count", "mw_fft2d", "mw_gerstner_from_material", "mw_gerstner_append_level_one",
    "mw_gerstner_displace", "mw_renderer_create", "mw_renderer_destroy", "mw_renderer_render_initial",
    "mw_renderer_set_initial", "mw_renderer_get_initial", "mw_renderer_set_phase", "mw_renderer_get_phase",
    "mw_renderer_set_params", "mw_renderer_generate_texture", "mw_renderer_sync", "mw_mesh_generate", "mw_wave_displace",
    "mw_tiles_create", "mw_tiles_destroy", "mw_tiles_disconnect", "mw_tiles_get_layout", "mw_tiles_export", "mw_tiles_connect",
    "mw_tiles_init_spectrum", "mw_tiles_set_h0", "mw_tiles_set_stream", "mw_tiles_generate_allgather", "mw_tiles_generate_local",
    "mw_tiles_allgather", "mw_tiles_wait", "mw_tiles_sync", "mw_tiles_gather_impl", "mw_tiles_ocean",
)
MW_TILES_MAX_WORLD = 16
MW_TILES_BLOB_BYTES = 512
MW_GATHER_NCCL, MW_GATHER_PEER, MW_GATHER_AUTO = 0, 1, 2
MW_TILES_ASYNC = 1 << 0
MW_TILES_PUSH_CE, MW_TILES_PUSH_SM, MW_TILES_PUSH_TMA = 1 << 1, 1 << 2, 1 << 3


class MwError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mistral_ocean error {code}: {message}")
        self.code = code
        self.message = message


class OceanParams(C.Structure):
    _fields_ = [
        ("resolution", C.c_int32), ("unit_width", C.c_float), ("length", C.c_float), ("choppiness", C.c_float),
        ("amplitude", C.c_float), ("wind_x", C.c_float), ("wind_y", C.c_float), ("t_division", C.c_float),
        ("seed", C.c_uint64), ("device", C.c_int32), ("tiles", C.c_int32), ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class TilesParams(C.Structure):
    _fields_ = [("ocean", OceanParams), ("world", C.c_int32), ("rank", C.c_int32), ("tiles_per_rank", C.c_int32),
                ("gather", C.c_int32), ("devices", C.c_int32 * 16), ("wind_step_deg", C.c_float), ("flags", C.c_uint32)]


class TilesLayout(C.Structure):
    _fields_ = [("slot_floats", C.c_int64), ("height_off", C.c_int64), ("disp_off", C.c_int64), ("normal_off", C.c_int64),
                ("whitecap_off", C.c_int64), ("world", C.c_int32), ("tiles_per_rank", C.c_int32), ("resolution", C.c_int32),
                ("local_ranks", C.c_int32)]


class OceanOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("height", "disp", "normal", "whitecap", "jacobian", "vertices", "colors")]


class RendererParams(C.Structure):
    _fields_ = [
        ("resolution", C.c_int32), ("unit_width", C.c_float), ("length", C.c_float), ("choppiness", C.c_float),
        ("amplitude", C.c_float), ("wind_x", C.c_float), ("wind_y", C.c_float), ("mult", C.c_float),
        ("seed1", C.c_float), ("seed2", C.c_float), ("device", C.c_int32), ("tiles", C.c_int32), ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class RendererOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("displacement", "height", "normal", "white", "white_rgba", "jacobian")]


class WaveParams(C.Structure):
    _fields_ = [("amplitude", C.c_float), ("frequency", C.c_float), ("speed", C.c_float), ("smoothing", C.c_float),
                ("device", C.c_int32), ("flags", C.c_uint32)]


class GerstnerWave(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("dir_x", "dir_y", "freq", "rate", "amp_xz", "amp_y")]


class GerstnerParams(C.Structure):
    _fields_ = [("n_waves", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("smoothing", C.c_float),
                ("waves", GerstnerWave * MW_GERSTNER_MAX_WAVES)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, fp = C.c_void_p, C.c_void_p  # buffers go through as raw addresses (host or device)
    lib.mw_version.restype = C.c_int
    lib.mw_last_error.restype = C.c_char_p
    lib.mw_kernel_launch_count.restype = C.c_int64
    lib.mw_ocean_create.argtypes = [C.POINTER(OceanParams), C.POINTER(vp)]
    lib.mw_ocean_destroy.argtypes = [vp]
    lib.mw_ocean_destroy.restype = None
    lib.mw_ocean_init_spectrum.argtypes = [vp]
    lib.mw_ocean_set_h0.argtypes = [vp, fp, fp]
    lib.mw_ocean_get_h0.argtypes = [vp, fp, fp]
    lib.mw_ocean_get_rest_vertices.argtypes = [vp, fp]
    lib.mw_ocean_get_dispersion.argtypes = [vp, fp]
    lib.mw_ocean_evolve_spectrum.argtypes = [vp, C.c_float, fp]
    lib.mw_ocean_generate.argtypes = [vp, C.c_float, C.POINTER(OceanOut)]
    lib.mw_ocean_update.argtypes = [vp, C.c_float, C.POINTER(OceanOut)]
    lib.mw_ocean_reset_timer.argtypes = [vp]
    lib.mw_ocean_timer.argtypes = [vp]
    lib.mw_ocean_timer.restype = C.c_float
    lib.mw_ocean_sync.argtypes = [vp]
    lib.mw_ocean_set_stream.argtypes = [vp, vp]
    lib.mw_ocean_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_int]
    lib.mw_fft2d.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int, fp, fp]
    lib.mw_gerstner_from_material.argtypes = [C.POINTER(GerstnerParams), C.c_float, C.c_float, C.c_float,
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.mw_gerstner_append_level_one.argtypes = [C.POINTER(GerstnerParams), C.c_float, C.c_float, C.c_float]
    lib.mw_gerstner_displace.argtypes = [C.POINTER(GerstnerParams), fp, fp, fp, C.c_int64, C.c_float, vp]
    lib.mw_renderer_create.argtypes = [C.POINTER(RendererParams), C.POINTER(vp)]
    lib.mw_renderer_destroy.argtypes = [vp]
    lib.mw_renderer_destroy.restype = None
    lib.mw_renderer_render_initial.argtypes = [vp]
    lib.mw_renderer_set_initial.argtypes = [vp, fp]
    lib.mw_renderer_get_initial.argtypes = [vp, fp]
    lib.mw_renderer_set_phase.argtypes = [vp, fp]
    lib.mw_renderer_get_phase.argtypes = [vp, fp]
    lib.mw_renderer_set_params.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    lib.mw_renderer_generate_texture.argtypes = [vp, C.c_float, C.POINTER(RendererOut)]
    lib.mw_renderer_sync.argtypes = [vp]
    lib.mw_mesh_generate.argtypes = [C.c_int, C.c_int32, C.c_float, fp, fp, fp, fp]
    lib.mw_wave_displace.argtypes = [C.POINTER(WaveParams), fp, fp, fp, C.c_int64, C.c_float, vp]
    lib.mw_tiles_create.argtypes = [C.POINTER(TilesParams), C.POINTER(vp)]
    lib.mw_tiles_destroy.argtypes = [vp]
    lib.mw_tiles_destroy.restype = None
    lib.mw_tiles_disconnect.argtypes = [vp]
    lib.mw_tiles_get_layout.argtypes = [vp, C.POINTER(TilesLayout)]
    lib.mw_tiles_export.argtypes = [vp, vp]
    lib.mw_tiles_connect.argtypes = [vp, vp]
    lib.mw_tiles_init_spectrum.argtypes = [vp]
    lib.mw_tiles_set_h0.argtypes = [vp, C.c_int, fp, fp]
    lib.mw_tiles_set_stream.argtypes = [vp, C.POINTER(vp)]
    lib.mw_tiles_generate_allgather.argtypes = [vp, C.c_float, C.POINTER(vp)]
    lib.mw_tiles_generate_local.argtypes = [vp, C.c_float, C.POINTER(vp)]
    lib.mw_tiles_allgather.argtypes = [vp]
    lib.mw_tiles_wait.argtypes = [vp, C.c_int]
    lib.mw_tiles_sync.argtypes = [vp]
    lib.mw_tiles_gather_impl.argtypes = [vp]
    lib.mw_tiles_ocean.argtypes = [vp, C.c_int]
    lib.mw_tiles_ocean.restype = vp
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != MW_OK:
        raise MwError(rc, load().mw_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().mw_kernel_launch_count())
