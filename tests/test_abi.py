"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/mistral_ocean.h declares, its structs have the layout the header states, and it fails loudly
(status code + message, no crash, no fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "mistral_ocean.h")).read()


def test_library_exports_every_declared_symbol(mw):
    declared = set(re.findall(r"\b(mw_[a-z0-9_]+)\s*\(", _header()))
    assert declared, "header parse failed"
    lib = mw.native.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mistral_ocean.h but not exported"
    assert declared == set(mw.native.EXPORTS)


def test_version_and_error_text(mw):
    lib = mw.native.load()
    assert lib.mw_version() == int(re.search(r"#define MW_VERSION (\d+)", _header()).group(1))
    assert isinstance(lib.mw_last_error(), bytes)


def test_struct_layouts_match_header(mw):
    n = mw.native
    assert C.sizeof(n.OceanParams) == 56 and n.OceanParams.seed.offset == 32
    assert C.sizeof(n.OceanOut) == 7 * 8
    assert C.sizeof(n.GerstnerWave) == 24
    assert C.sizeof(n.GerstnerParams) == 16 + 64 * 24
    assert C.sizeof(n.RendererParams) == 56 and n.RendererParams.seed1.offset == 32 and n.RendererParams.flags.offset == 48
    assert C.sizeof(n.RendererOut) == 6 * 8 and C.sizeof(n.WaveParams) == 24
    fields = re.search(r"typedef struct mw_ocean_params \{(.*?)\} mw_ocean_params;", _header(), re.S).group(1)
    names = re.findall(r"\b(?:int32_t|uint32_t|uint64_t|float)\s+([a-z_0-9]+);", fields)
    assert names == [f[0] for f in n.OceanParams._fields_]
    out_fields = re.search(r"typedef struct mw_ocean_out \{(.*?)\} mw_ocean_out;", _header(), re.S).group(1)
    assert re.findall(r"float\*\s+([a-z]+);", out_fields) == [f[0] for f in n.OceanOut._fields_]


@pytest.mark.parametrize("kw,frag", [
    (dict(resolution=300), "power of two"),            # not a power of two and too large for the direct-sum kernel
    (dict(resolution=1), "power of two"),
    (dict(resolution=4096), "power of two"),
    (dict(resolution=512, length=500.0), "periodic"),  # length != resolution * unit_width above the direct-sum limit
    (dict(resolution=64, unit_width=-1.0, length=-64.0), "positive"),
    (dict(resolution=64, tiles=0), "tiles"),
    (dict(resolution=64, t_division=0.0), "t_division"),
])
def test_create_validates_before_touching_the_gpu(mw, kw, frag):
    with pytest.raises(mw.native.MwError) as ei:
        mw.Ocean(**kw)
    assert ei.value.code == mw.native.MW_E_INVALID_ARG
    assert frag in ei.value.message


def test_null_arguments_are_errors_not_crashes(mw):
    lib = mw.native.load()
    assert lib.mw_ocean_create(None, None) == mw.native.MW_E_INVALID_ARG
    assert lib.mw_ocean_generate(None, 0.0, None) == mw.native.MW_E_INVALID_ARG
    assert lib.mw_ocean_init_spectrum(None) == mw.native.MW_E_INVALID_ARG
    assert lib.mw_fft2d(0, 64, 1, -1, None, None) == mw.native.MW_E_INVALID_ARG
    x = np.zeros((48, 48), np.complex64)
    assert lib.mw_fft2d(0, 48, 1, -1, x.ctypes.data, x.ctypes.data) == mw.native.MW_E_INVALID_ARG
    assert lib.mw_gerstner_displace(None, None, None, None, 0, 0.0, None) == mw.native.MW_E_INVALID_ARG
    lib.mw_ocean_destroy(None)  # no-op
    # multi-GPU tile sets
    n = mw.native
    h = C.c_void_p()
    assert lib.mw_tiles_create(None, C.byref(h)) == n.MW_E_INVALID_ARG
    p = n.TilesParams()
    p.ocean = n.OceanParams(64, 1.0, 64.0, 1.0, 0.01, 5.0, 3.0, 1.0, 1000, 0, 1, 0, 0)
    p.world, p.rank, p.tiles_per_rank, p.gather = 99, -1, 1, n.MW_GATHER_PEER
    assert lib.mw_tiles_create(C.byref(p), C.byref(h)) == n.MW_E_INVALID_ARG and b"world" in lib.mw_last_error()
    p.world, p.rank = 2, 2
    assert lib.mw_tiles_create(C.byref(p), C.byref(h)) == n.MW_E_INVALID_ARG and b"rank" in lib.mw_last_error()
    p.rank, p.gather = 0, 7
    assert lib.mw_tiles_create(C.byref(p), C.byref(h)) == n.MW_E_INVALID_ARG and b"gather" in lib.mw_last_error()
    assert lib.mw_tiles_generate_allgather(None, 0.0, None) == n.MW_E_INVALID_ARG
    assert lib.mw_tiles_export(None, None) == n.MW_E_INVALID_ARG
    assert lib.mw_tiles_sync(None) == n.MW_E_INVALID_ARG
    assert not lib.mw_tiles_ocean(None, 0)
    lib.mw_tiles_destroy(None)  # no-op


def test_tiles_struct_layouts_match_header(mw):
    n = mw.native
    assert C.sizeof(n.TilesParams) == 56 + 4 * 4 + 16 * 4 + 8 and n.TilesParams.devices.offset == 72
    assert C.sizeof(n.TilesLayout) == 5 * 8 + 4 * 4
    hdr = _header()
    assert int(re.search(r"#define MW_TILES_BLOB_BYTES (\d+)", hdr).group(1)) == n.MW_TILES_BLOB_BYTES
    assert int(re.search(r"#define MW_TILES_MAX_WORLD (\d+)", hdr).group(1)) == n.MW_TILES_MAX_WORLD


def _header_enumerators():
    """name -> value of every `MW_X = <int or 1u << k>` enumerator in the header"""
    out = {}
    for name, val in re.findall(r"\b(MW_[A-Z0-9_]+)\s*=\s*(-?\d+u?\s*(?:<<\s*\d+)?)", _header()):
        m = re.fullmatch(r"(-?\d+)u?\s*(?:<<\s*(\d+))?", val.strip())
        out[name] = int(m.group(1)) << int(m.group(2) or 0)
    return out


def test_enumerators_match_header_in_python_and_csharp(mw):
    """Every status code / flag / gather mode of include/mistral_ocean.h carries the same value in the ctypes mirror, and the
    ones the C# binding spells out agree too (bindings/MistralOceanNative.cs is source only: nothing here compiles it)."""
    enums = _header_enumerators()
    for must in ("MW_E_NCCL", "MW_DEVICE_PTRS", "MW_GATHER_AUTO", "MW_TILES_ASYNC", "MW_TILES_PUSH_CE", "MW_TILES_PUSH_SM", "MW_TILES_PUSH_TMA"):
        assert must in enums, must
    n = mw.native
    for name, val in enums.items():
        assert hasattr(n, name), f"{name} missing in mistral-water_b200/native.py"
        assert getattr(n, name) == val, (name, getattr(n, name), val)
    cs = open(os.path.join(ROOT, "bindings", "MistralOceanNative.cs")).read()
    spelled = dict((k, int(v)) for k, v in re.findall(r"\b(MW_[A-Z0-9_]+)\s*=\s*(-?\d+)\b", cs))
    assert {"MW_GATHER_AUTO", "MW_TILES_PUSH_TMA"} <= set(spelled)
    for name, val in spelled.items():
        if name in enums:
            assert enums[name] == val, (name, val, enums[name])


def test_no_cpu_fallback_without_a_device(mw):
    """On a box without a GPU the engine must refuse, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mw.native.MwError) as ei:
        mw.Ocean(64)
    assert ei.value.code == mw.native.MW_E_CUDA
    g = mw.pond_wave_table_32()
    with pytest.raises(mw.native.MwError):
        g.displace(np.zeros((8, 3), np.float32), 0.0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or the C ABI may reference it."""
    pkg = os.path.join(ROOT, "mistral-water_b200")
    for dp, _, fns in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("oracle/ is", ""), f"{fn} mentions the oracle"
                assert "import cref" not in txt and "ref_fft64" not in txt


def test_csharp_binding_declares_every_symbol_and_layout(mw):
    """bindings/MistralOceanNative.cs cannot be compiled here (no dotnet); what can be checked is that it declares every
    exported symbol and that LayoutCheck.cs expects the sizes the ctypes mirror has."""
    n = mw.native
    cs = open(os.path.join(ROOT, "bindings", "MistralOceanNative.cs")).read()
    declared = set(re.findall(r"static extern \w+ (mw_[a-z0-9_]+)\(", cs))
    assert declared == set(n.EXPORTS), sorted(set(n.EXPORTS) ^ declared)
    chk = open(os.path.join(ROOT, "bindings", "LayoutCheck.cs")).read()
    want = {"MwOceanParams": n.OceanParams, "MwOceanOut": n.OceanOut, "MwGerstnerWave": n.GerstnerWave,
            "MwGerstnerParams": n.GerstnerParams, "MwRendererParams": n.RendererParams, "MwRendererOut": n.RendererOut,
            "MwWaveParams": n.WaveParams, "MwTilesParams": n.TilesParams, "MwTilesLayout": n.TilesLayout}
    for name, ct in want.items():
        m = re.search(r"Marshal\.SizeOf\(typeof\(%s\)\), ([0-9 +*]+)\)" % name, chk)
        assert m, name
        assert eval(m.group(1)) == C.sizeof(ct), name
    proj = open(os.path.join(ROOT, "bindings", "MistralOcean.Bindings.csproj")).read()
    for f in ("MistralOceanNative.cs", "UnityShim.cs", "LayoutCheck.cs"):
        assert f in proj
