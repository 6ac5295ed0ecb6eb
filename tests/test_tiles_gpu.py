"""The multi-GPU tile set through the C ABI (mw_tiles_*, include/mistral_ocean.h), checked against single-handle runs of
the same tiles -- not against itself.  One-GPU boxes run the world = 1 cases; the rest need 2 (or 4) GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, TPR = 128, 2


def _expected(mw, torch, world, times, device=0):
    from mistral_water_b200.tiles import FIELDS, tile_wind
    dev = torch.device("cuda", device)
    exp = {}
    for r in range(world):
        o = mw.Ocean(N, seed=1000 + r * TPR, tiles=TPR, wind=tile_wind((5.0, 3.0), r * TPR), device=device, device_ptrs=True)
        o.init_spectrum()
        for t in times:
            bufs = {k: torch.empty(TPR * N * N * c, device=dev) for k, c in FIELDS}
            o.generate(t, bufs)
            o.sync()
            exp[(t, r)] = torch.cat([bufs[k] for k, _ in FIELDS]).cpu()
        o.close()
    return exp


@pytest.mark.gpu
@pytest.mark.parametrize("gather", ["peer", "nccl"])
def test_world_one_tile_set_equals_single_handle(mw, gather):
    import torch
    from mistral_water_b200.tiles import TileSet
    times = [0.0, 1.7, 60.0]
    exp = _expected(mw, torch, 1, times)
    with TileSet(N, 1, rank=None, tiles_per_rank=TPR, gather=gather, asynchronous=False) as ts:
        assert ts.slot_floats == TPR * N * N * 7 and ts.local_ranks == 1
        assert ts.field_off == {"height": 0, "disp": TPR * N * N, "normal": 3 * TPR * N * N, "whitecap": 6 * TPR * N * N}
        ts.init_spectrum()
        for t in times:
            (ptr,) = ts.generate_allgather(t)
            assert torch.equal(ts.as_tensor(ptr)[0].cpu(), exp[(t, 0)])
        # the two halves, and the borrowed ocean handle
        (ptr,) = ts.generate_local(1.7)
        ts.allgather(); ts.sync()
        assert torch.equal(ts.as_tensor(ptr)[0].cpu(), exp[(1.7, 0)])
        assert ts.ocean_handle(0) != 0 and ts.ocean_handle(1) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("gather,push", [("peer", "tma"), ("peer", "sm"), ("peer", "ce"), ("nccl", "tma")])
def test_single_process_drives_two_gpus(mw, gather, push):
    """ONE process, no torch.distributed: what a single C# host does (ncclCommInitAll / peer pushes fenced by events), with each
    of the peer arm's push engines (TMA bulk-copy kernel, SM-store kernel, copy engines)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mistral_water_b200.tiles import TileSet
    world = 2
    times = [0.0, 1.7, 3.25, 60.0]
    exp = _expected(mw, torch, world, times)
    with TileSet(N, world, rank=None, devices=[0, 1], tiles_per_rank=TPR, gather=gather, asynchronous=True, push=push) as ts:
        assert ts.local_ranks == world and ts.gather_impl == gather
        ts.init_spectrum()
        # back-to-back frames with no host synchronisation: frame k's gather overlaps frame k + 1's generation
        ptrs = [ts.generate_allgather(t) for t in times]
        ts.sync()
        for t, pp in zip(times[-2:], ptrs[-2:]):        # the two live buffers
            for i in range(world):                       # every device's copy holds every rank's slot
                g = ts.as_tensor(pp[i], i).cpu()
                for r in range(world):
                    assert torch.equal(g[r], exp[(t, r)]), (gather, t, i, r)
        assert ptrs[0] == ptrs[2] and ptrs[1] == ptrs[3] and ptrs[0] != ptrs[1]   # double buffering


def _torchrun(world, extra_env=None):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "tiles_check.py")],
                          capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)


@pytest.mark.gpu
@pytest.mark.parametrize("world,push", [(2, "tma"), (2, "sm"), (2, "ce"), (4, "tma")])
def test_one_process_per_gpu_gathers_match_single_handles(world, push):
    """torchrun: every gathered tile bit-equal to a single-handle run, for the peer arm (with the push engine named; the
    MW_TILES_PUSH override reaches mw_tiles_create in every rank) and for the ncclAllGather arm."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, {"MW_TILES_PUSH": push})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "TILES_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_tiles_connect_rejects_foreign_blobs(mw):
    """A rank handle stays unconnected (MW_E_STATE on use) until it sees consistent blobs."""
    import ctypes as C
    n = mw.native
    lib = n.load()
    p = n.TilesParams()
    p.ocean = n.OceanParams(64, 1.0, 64.0, 1.0, 0.01, 5.0, 3.0, 1.0, 1000, 0, 1, 0, 0)
    p.world, p.rank, p.tiles_per_rank, p.gather = 2, 0, 1, n.MW_GATHER_PEER
    h = C.c_void_p()
    n.check(lib.mw_tiles_create(C.byref(p), C.byref(h)))
    try:
        assert lib.mw_tiles_generate_allgather(h, 0.0, None) == n.MW_E_STATE
        blob = C.create_string_buffer(n.MW_TILES_BLOB_BYTES)
        n.check(lib.mw_tiles_export(h, blob))
        both = C.create_string_buffer(blob.raw + blob.raw, 2 * n.MW_TILES_BLOB_BYTES)   # rank 1's blob is not rank 1's
        assert lib.mw_tiles_connect(h, both) == n.MW_E_INVALID_ARG
        assert b"blob 1" in lib.mw_last_error()
    finally:
        lib.mw_tiles_destroy(h)
