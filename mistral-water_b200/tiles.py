"""Independent ocean tiles sharded one-per-GPU, with ONE all-gather of the final float buffers
(BASELINE config 5).

Rank r owns tiles [r * tpr, (r + 1) * tpr).  Tiles never exchange data while being generated (nothing
in FFTMesh.cs couples two meshes), so the only collective is the final in-place all-gather.  The whole
mechanism -- gather buffers, streams, double buffering, fences, NCCL communicators, peer mappings --
lives behind the C ABI (mw_tiles_*, csrc/mw_tiles.cu); this module is the ctypes caller:

  TileSet       one mw_tiles handle: all GPUs from ONE process (rank=None; the reference's host model,
                one Unity process) or one process per GPU (rank=r; blobs swapped by `exchange`)
  ShardedTiles  the torchrun-shaped convenience around it (blobs travel through torch.distributed);
                with `make_generator` it runs a CPU stub instead (gloo tests of layout / sharding)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

FIELDS = (("height", 1), ("disp", 2), ("normal", 3), ("whitecap", 1))  # 7 floats = 28 B per grid point
FLOATS_PER_POINT = sum(c for _, c in FIELDS)


@dataclass(frozen=True)
class TileLayout:
    """Where each field of each rank lives inside the gather buffer [world][slot_floats]."""
    N: int
    world: int
    tiles_per_rank: int = 1

    @property
    def points_per_rank(self) -> int:
        return self.tiles_per_rank * self.N * self.N

    @property
    def slot_floats(self) -> int:
        return self.points_per_rank * FLOATS_PER_POINT

    @property
    def slot_bytes(self) -> int:
        return self.slot_floats * 4

    def field_range(self, name: str) -> tuple[int, int]:
        """[begin, end) float offsets of a field inside one slot: fields are planar, [tile][idx][comp]."""
        off = 0
        for f, c in FIELDS:
            n = self.points_per_rank * c
            if f == name:
                return off, off + n
            off += n
        raise KeyError(name)

    def global_tile(self, rank: int, local: int) -> int:
        return rank * self.tiles_per_rank + local

    def owner(self, global_tile: int) -> tuple[int, int]:
        return divmod(global_tile, self.tiles_per_rank)


def tile_wind(base_wind, global_tile: int, step_deg: float = 45.0):
    """Config 5: tile k's wind is the base wind rotated by 45 deg * k."""
    a = (step_deg * global_tile) * math.pi / 180.0   # same operation order as csrc/mw_tiles.cu (rotate_wind)
    c, s = math.cos(a), math.sin(a)
    return (c * base_wind[0] - s * base_wind[1], s * base_wind[0] + c * base_wind[1])


class _DevArray:
    """A device allocation owned by the library, presented through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class TileSet:
    """One mw_tiles handle (include/mistral_ocean.h, "Multi-GPU tile sets").  Every method is one C-ABI call.

    rank=None : this process drives all `world` GPUs (`devices`, default 0..world-1); no blob exchange, no
                torch.distributed -- what a single C# host process does.
    rank=r    : one process per GPU; `exchange(blob) -> [blob of rank 0, blob of rank 1, ...]` is any all-gather of
                512-byte strings the host has (ShardedTiles passes torch.distributed.all_gather_object).
    gather    : "auto" (the peer pushes; include/mistral_ocean.h MW_GATHER_AUTO), "peer", "nccl".
    push      : what moves a slot into the peers' buffers in the peer arm -- "tma" (bulk-copy kernel, default), "sm" (store kernel),
                "ce" (copy engines): the MW_TILES_PUSH_* flags.
    """

    def __init__(self, N: int, world: int, rank: int | None = None, devices=None, tiles_per_rank: int = 1,
                 gather: str = "auto", base_seed: int = 1000, wind=(5.0, 3.0), amplitude: float = 0.01,
                 unit_width: float = 1.0, choppiness: float = 1.0, wind_step_deg: float = 45.0, asynchronous: bool = True,
                 profile: bool = False, exchange=None, push: str = "tma"):
        import ctypes as C

        import numpy as np

        from . import native

        self._C, self._native = C, native
        self._lib = native.load()
        self.N, self.world, self.rank, self.tiles_per_rank = int(N), int(world), rank, int(tiles_per_rank)
        devices = list(range(world)) if devices is None else list(devices)
        p = native.TilesParams()
        length = float(np.float32(N) * np.float32(unit_width))
        p.ocean = native.OceanParams(int(N), float(unit_width), length, float(choppiness), float(amplitude), float(wind[0]),
                                     float(wind[1]), 1.0, int(base_seed), 0, 1, native.MW_PROFILE if profile else 0, 0)
        p.world, p.rank, p.tiles_per_rank = self.world, (-1 if rank is None else int(rank)), self.tiles_per_rank
        p.gather = {"nccl": native.MW_GATHER_NCCL, "peer": native.MW_GATHER_PEER, "p2p": native.MW_GATHER_PEER,
                    "auto": native.MW_GATHER_AUTO}[gather]
        for i, d in enumerate(devices[:native.MW_TILES_MAX_WORLD]):
            p.devices[i] = int(d)
        p.wind_step_deg = float(wind_step_deg)
        p.flags = (native.MW_TILES_ASYNC if asynchronous else 0) | {"ce": native.MW_TILES_PUSH_CE, "sm": native.MW_TILES_PUSH_SM,
                                                                  "tma": native.MW_TILES_PUSH_TMA}[push]
        self.devices = devices
        self._h = C.c_void_p()
        native.check(self._lib.mw_tiles_create(C.byref(p), C.byref(self._h)))
        lay = native.TilesLayout()
        native.check(self._lib.mw_tiles_get_layout(self._h, C.byref(lay)))
        self.slot_floats, self.local_ranks = int(lay.slot_floats), int(lay.local_ranks)
        self.field_off = {"height": int(lay.height_off), "disp": int(lay.disp_off), "normal": int(lay.normal_off),
                          "whitecap": int(lay.whitecap_off)}
        if rank is not None and world > 1:
            if exchange is None:
                raise ValueError("one process per GPU needs `exchange` to swap the ranks' blobs")
            blob = C.create_string_buffer(native.MW_TILES_BLOB_BYTES)
            native.check(self._lib.mw_tiles_export(self._h, blob))
            blobs = exchange(blob.raw)
            if len(blobs) != world or any(len(b) != native.MW_TILES_BLOB_BYTES for b in blobs):
                raise ValueError("exchange() must return one 512-byte blob per rank, in rank order")
            allb = C.create_string_buffer(b"".join(blobs), native.MW_TILES_BLOB_BYTES * world)
            native.check(self._lib.mw_tiles_connect(self._h, allb))
        self.gather_impl = {native.MW_GATHER_NCCL: "nccl", native.MW_GATHER_PEER: "peer"}[int(self._lib.mw_tiles_gather_impl(self._h))]

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mw_tiles_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- calls
    def _ptrs(self):
        return (self._C.c_void_p * self.local_ranks)()

    def init_spectrum(self) -> None:
        self._native.check(self._lib.mw_tiles_init_spectrum(self._h))

    def set_h0(self, local_rank: int, h0_ptr: int, h0conj_ptr: int) -> None:
        """Device pointers on that rank's device; read in user-stream order."""
        self._native.check(self._lib.mw_tiles_set_h0(self._h, int(local_rank), self._C.c_void_p(h0_ptr), self._C.c_void_p(h0conj_ptr)))

    def set_stream(self, streams) -> None:
        """streams: one cudaStream_t (int) per local rank, or None for the handle's own."""
        arr = None
        if streams is not None:
            arr = self._ptrs()
            for i, s in enumerate(streams):
                arr[i] = int(s) if s else None
        self._native.check(self._lib.mw_tiles_set_stream(self._h, arr))

    def generate_allgather(self, t: float) -> list[int]:
        out = self._ptrs()
        self._native.check(self._lib.mw_tiles_generate_allgather(self._h, float(t), out))
        return [int(x) for x in out]

    def generate_local(self, t: float) -> list[int]:
        out = self._ptrs()
        self._native.check(self._lib.mw_tiles_generate_local(self._h, float(t), out))
        return [int(x) for x in out]

    def allgather(self) -> None:
        self._native.check(self._lib.mw_tiles_allgather(self._h))

    def wait(self, frames_back: int = 0) -> None:
        self._native.check(self._lib.mw_tiles_wait(self._h, int(frames_back)))

    def sync(self) -> None:
        self._native.check(self._lib.mw_tiles_sync(self._h))

    def disconnect(self) -> None:
        """Drop the peer mappings / NCCL communicator (then: barrier across ranks, then close())."""
        self._native.check(self._lib.mw_tiles_disconnect(self._h))

    def ocean_handle(self, local_rank: int = 0) -> int:
        return int(self._lib.mw_tiles_ocean(self._h, int(local_rank)) or 0)

    def as_tensor(self, ptr: int, local_rank: int = 0):
        """[world][slot_floats] torch view of a gather buffer returned by generate_*."""
        import torch

        dev = self.devices[local_rank if self.rank is None else self.rank]
        return torch.as_tensor(_DevArray(ptr, self.world * self.slot_floats), device=torch.device("cuda", dev)).view(self.world, self.slot_floats)


class ShardedTiles:
    """This rank's share of the tile set under torchrun (one process per GPU).

    CUDA: a TileSet whose blobs travel through torch.distributed; `stream` (a torch stream) is the user stream the
    calls are ordered against.  `make_generator(rank_params) -> callable(t, slot_views)` replaces the engine by a
    stub for the CPU `gloo` tests of the sharding / layout / collective logic.
    """

    def __init__(self, N: int, rank: int, world: int, tiles_per_rank: int = 1, base_seed: int = 1000,
                 wind=(5.0, 3.0), amplitude: float = 0.01, unit_width: float = 1.0, choppiness: float = 1.0,
                 device=None, group=None, make_generator=None, gather: str | None = None, profile: bool = False):
        import torch

        self.torch = torch
        self.layout = TileLayout(N, world, tiles_per_rank)
        self.rank, self.world, self.group = rank, world, group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.rank_params = dict(resolution=N, unit_width=unit_width, choppiness=choppiness, amplitude=amplitude,
                                wind=tile_wind(wind, self.layout.global_tile(rank, 0)),
                                seed=base_seed + self.layout.global_tile(rank, 0), tiles=tiles_per_rank)
        self.tileset = None
        if make_generator is not None or self.device.type != "cuda":
            # CPU stub path: one gather buffer, the collective through the process group (gloo)
            if make_generator is None:
                raise RuntimeError("ShardedTiles on a CPU device needs make_generator: the engine has no CPU path")
            self.gather_impl = "nccl"
            self.gather = torch.empty((world, self.layout.slot_floats), dtype=torch.float32, device=self.device)
            self._gen = make_generator(self.rank_params)
            return
        import os

        import torch.distributed as dist

        def exchange(blob: bytes):
            out = [None] * world
            dist.all_gather_object(out, blob, group=group)
            return out

        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        devices = [0] * world
        devices[rank] = dev_index
        self.tileset = TileSet(N, world, rank=rank, devices=devices, tiles_per_rank=tiles_per_rank,
                               gather=gather or os.environ.get("MW_GATHER", "auto"), base_seed=base_seed, wind=wind,
                               amplitude=amplitude, unit_width=unit_width, choppiness=choppiness, asynchronous=True,
                               profile=profile, exchange=exchange if world > 1 else None)
        self.gather_impl = self.tileset.gather_impl
        self.stream = torch.cuda.Stream(device=self.device)
        self.tileset.set_stream([self.stream.cuda_stream])
        self.tileset.init_spectrum()
        self.tileset.sync()
        self.gather = None

    # ------------------------------------------------------------------ views
    def slot_views(self, rank: int | None = None) -> dict:
        """Field views into one rank's slot (default: ours) of the current gather buffer."""
        slot = self.gather[self.rank if rank is None else rank]
        out = {}
        for name, comps in FIELDS:
            b, e = self.layout.field_range(name)
            out[name] = slot[b:e]
        return out

    def tile_view(self, global_tile: int, name: str):
        """Field `name` of any tile, from the gathered buffer: [N*N, comps]."""
        r, l = self.layout.owner(global_tile)
        b, e = self.layout.field_range(name)
        comps = dict(FIELDS)[name]
        n2 = self.layout.N * self.layout.N
        return self.gather[r, b:e].view(self.layout.tiles_per_rank, n2, comps)[l]

    # ------------------------------------------------------------------ frames
    def _view(self, ptrs):
        self.gather = self.tileset.as_tensor(ptrs[0])
        return self.gather

    def generate_local(self, t: float):
        """Produce this rank's tiles into its slot (asynchronous; no collective)."""
        if self.tileset is None:
            self._gen(float(t), self.slot_views())
            return self.gather
        return self._view(self.tileset.generate_local(t))

    def all_gather(self) -> None:
        """The one collective: in-place all-gather of the slots of the frame last generated."""
        if self.world == 1:
            return
        if self.tileset is None:
            import torch.distributed as dist

            dist.all_gather_into_tensor(self.gather.view(-1), self.gather[self.rank].view(-1).clone(), group=self.group)
            return
        self.tileset.allgather()

    def generate(self, t: float):
        """One frame + its all-gather; the returned buffer is complete once `stream` (the user stream) gets there."""
        if self.tileset is None:
            self.generate_local(t)
            self.all_gather()
            return self.gather
        g = self._view(self.tileset.generate_allgather(t))
        self.tileset.wait(0)
        return g

    def generate_pipelined(self, t: float):
        """Frame k: generated into gather buffer k % 2 and gathered on the library's communication streams, so that the
        collective of frame k runs under the generation of frame k + 1.  Returns the buffer being gathered; call
        finish() before reading it (on `stream`)."""
        if self.tileset is None:
            return self.generate(t)
        return self._view(self.tileset.generate_allgather(t))

    def finish(self) -> None:
        """Make the user stream wait for everything generate_pipelined() queued."""
        if self.tileset is not None:
            self.tileset.wait(0)

    def sync(self) -> None:
        if self.tileset is not None:
            self.tileset.sync()

    def close(self) -> None:
        """Collective at world > 1: every rank drops its peer mappings, the ranks meet, then the buffers are freed (an
        exported allocation must not be freed while a peer still maps it)."""
        if self.tileset is not None:
            self.tileset.sync()
            if self.world > 1:
                import torch.distributed as dist

                self.tileset.disconnect()
                dist.barrier(group=self.group)
            self.tileset.close()
            self.tileset = None
