"""Independent ocean tiles sharded one-per-GPU, with ONE all-gather of the final float buffers
(BASELINE config 5).

One process per GPU (torchrun); rank r owns tiles [r * tpr, (r + 1) * tpr).  Tiles never exchange
data while being generated (nothing in FFTMesh.cs couples two meshes), so the only collective is
the final in-place all-gather: every rank's engine writes its outputs straight into its own slot of
the gather buffer (device pointers handed to mw_ocean_generate), then
`all_gather_into_tensor(gather, gather[rank])` runs over NCCL / NVLink.

torch is used for what it is good at here -- device memory, streams, the process group.
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
    a = math.radians(step_deg * global_tile)
    c, s = math.cos(a), math.sin(a)
    return (c * base_wind[0] - s * base_wind[1], s * base_wind[0] + c * base_wind[1])


class ShardedTiles:
    """This rank's share of the tile set + the gather buffer.

    `make_generator(rank_params) -> callable(t, slot_views: dict[str, Tensor])` builds the local
    producer; the default is the CUDA engine (mistral_water_b200.Ocean with device pointers).  The
    CPU `gloo` tests inject a stub there to exercise the sharding / layout / collective logic.
    """

    def __init__(self, N: int, rank: int, world: int, tiles_per_rank: int = 1, base_seed: int = 1000,
                 wind=(5.0, 3.0), amplitude: float = 0.01, unit_width: float = 1.0, choppiness: float = 1.0,
                 device=None, group=None, make_generator=None):
        import torch

        self.torch = torch
        self.layout = TileLayout(N, world, tiles_per_rank)
        self.rank, self.world, self.group = rank, world, group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # two gather buffers: the all-gather of frame k (communication stream) overlaps the generation of
        # frame k + 1 (engine stream) -- see generate_pipelined()
        self.nbuf = 2 if world > 1 else 1
        self.gathers = [torch.empty((world, self.layout.slot_floats), dtype=torch.float32, device=self.device)
                        for _ in range(self.nbuf)]
        self.gather = self.gathers[0]
        self._frame = 0
        self._cur_buf = 0
        # how the one collective is carried out: "p2p" = every rank pushes its slot straight into the peers' gather
        # buffers over NVLink with the copy engines (CUDA IPC peer memory, no SMs taken from the frame kernels) and a
        # 4-byte NCCL all-reduce closes the step; "nccl" = ncclAllGather.  MW_GATHER=nccl|p2p overrides.
        self.gather_impl, self.p2p_error = "nccl", None
        if world > 1 and self.device.type == "cuda" and make_generator is None:
            import os
            want = os.environ.get("MW_GATHER", "p2p")
            if want == "p2p":
                self._setup_p2p()
        self.rank_params = dict(resolution=N, unit_width=unit_width, choppiness=choppiness, amplitude=amplitude,
                                wind=tile_wind(wind, self.layout.global_tile(rank, 0)),
                                seed=base_seed + self.layout.global_tile(rank, 0), tiles=tiles_per_rank)
        self._gen = (make_generator or self._cuda_generator)(self.rank_params)

    def _cuda_generator(self, rp):
        from .ocean import Ocean

        torch = self.torch
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.stream = torch.cuda.Stream(device=self.device)
        self.ocean = Ocean(device=dev_index, device_ptrs=True, **rp)
        self.ocean.set_stream(self.stream.cuda_stream)
        self.ocean.init_spectrum()
        self.ocean.sync()

        def run(t, views):
            self.ocean.generate(t, views)

        return run

    # ------------------------------------------------------------------ peer-memory all-gather
    def _setup_p2p(self) -> None:
        """Open every peer's gather buffers from THIS rank's device through the library's CUDA IPC entry points
        (mw_peer_export / mw_peer_open: one process per GPU, all GPUs of the node visible, NVLink peer access).
        Any failure leaves gather_impl == "nccl" with the reason in p2p_error; the ranks agree on the outcome."""
        torch = self.torch
        import torch.distributed as dist
        from . import native

        ok, err = True, None
        self._peer_ptr, self._peer_bases = {}, []
        try:
            mine = [native.peer_export(g.data_ptr()) for g in self.gathers]   # (64-byte IPC handle, offset) per buffer
            allh = [None] * self.world
            dist.all_gather_object(allh, (self.device.index, mine), group=self.group)
            opened = {}
            for r, (dev_index, handles) in enumerate(allh):
                if r == self.rank:
                    continue
                if not torch.cuda.can_device_access_peer(self.device.index, dev_index):
                    raise RuntimeError(f"no peer access from cuda:{self.device.index} to cuda:{dev_index}")
                ptrs = []
                for handle, off in handles:
                    if handle not in opened:                                  # an allocation is opened once per process
                        opened[handle] = native.peer_open(self.device.index, handle)
                        self._peer_bases.append(opened[handle])
                    ptrs.append(opened[handle] + off)
                self._peer_ptr[r] = ptrs                                      # rank r's gather buffers, as seen from here
        except Exception as e:  # noqa: BLE001
            ok, err = False, repr(e)
        flag = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.gather_impl = "p2p"
            self._push_streams = {r: torch.cuda.Stream(device=self.device) for r in self._peer_ptr}
            self._push_done = {r: torch.cuda.Event() for r in self._peer_ptr}
            self._token = torch.zeros(1, device=self.device, dtype=torch.int32)
        else:
            self.p2p_error = err or "a peer could not open the buffers"
            self._close_peers()

    def _close_peers(self) -> None:
        from . import native
        for base in getattr(self, "_peer_bases", []):
            try:
                native.peer_close(self.device.index, base)
            except Exception:  # noqa: BLE001
                pass
        self._peer_bases, self._peer_ptr = [], {}

    def _push_slots(self, buf: int) -> None:
        """Enqueue the all-gather of gather buffer `buf` on the current stream (which already waits for this rank's
        slot to be complete and for this rank's readers of the buffer):
          1. a 4-byte all-reduce -- every rank has reached this point, so nobody still reads what is about to be overwritten;
          2. this rank's slot pushed into every peer's buffer (mw_peer_copy), one copy-engine stream per peer, peers
             visited in a rank-dependent order so that no destination is hit by everybody at once;
          3. a second 4-byte all-reduce -- when it completes here, every rank's pushes have landed."""
        torch = self.torch
        import torch.distributed as dist
        from . import native

        cur = torch.cuda.current_stream(self.device)
        dist.all_reduce(self._token, group=self.group)
        ready = torch.cuda.Event()
        ready.record(cur)
        slot_bytes = self.layout.slot_bytes
        src = self.gathers[buf][self.rank].data_ptr()
        for i in range(1, self.world):
            r = (self.rank + i) % self.world
            ps = self._push_streams[r]
            ps.wait_event(ready)
            native.peer_copy(self._peer_ptr[r][buf] + self.rank * slot_bytes, src, slot_bytes, ps.cuda_stream)
            self._push_done[r].record(ps)
            cur.wait_event(self._push_done[r])
        dist.all_reduce(self._token, group=self.group)

    def slot_views(self, rank: int | None = None, buf: int = 0) -> dict:
        """Field views into one rank's slot (default: ours) of gather buffer `buf`."""
        slot = self.gathers[buf][self.rank if rank is None else rank]
        out = {}
        for name, comps in FIELDS:
            b, e = self.layout.field_range(name)
            out[name] = slot[b:e]
        return out

    def generate_local(self, t: float) -> None:
        """Produce this rank's tiles into its slot (asynchronous on the engine's stream)."""
        self._gen(float(t), self.slot_views())

    def all_gather(self) -> None:
        """The one collective: in-place all-gather of the slots."""
        import torch.distributed as dist

        if self.world == 1:
            return
        if self.gather_impl == "p2p":
            self._push_slots(self._cur_buf)
            return
        dist.all_gather_into_tensor(self.gather.view(-1), self.gather[self.rank].view(-1), group=self.group)

    def generate(self, t: float):
        torch = self.torch
        self.generate_local(t)
        if hasattr(self, "stream"):
            torch.cuda.current_stream(self.device).wait_stream(self.stream)  # NCCL runs after the producer
        self.all_gather()
        return self.gather

    def generate_pipelined(self, t: float):
        """Frame k: generate into gather buffer k % 2 on the engine stream, then all-gather it on a separate
        communication stream, so that the collective of frame k runs under the generation of frame k + 1.
        Returns the buffer being gathered; call finish() before reading it."""
        torch = self.torch
        if self.world == 1 or not hasattr(self, "stream"):
            return self.generate(t)
        if not hasattr(self, "comm_stream"):
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self._ev_gen = [torch.cuda.Event() for _ in range(2)]
            self._ev_comm = [torch.cuda.Event() for _ in range(2)]
            self._comm_used = [False, False]
        b = self._frame & 1
        self._frame += 1
        self._cur_buf = b
        ev_user = torch.cuda.Event()
        ev_user.record(torch.cuda.current_stream(self.device))
        if self._comm_used[b]:
            self.stream.wait_event(self._ev_comm[b])      # buffer b is free again once its last gather is done
        self._gen(float(t), self.slot_views(buf=b))
        self._ev_gen[b].record(self.stream)
        import torch.distributed as dist
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(self._ev_gen[b])
            g = self.gathers[b]
            if self.gather_impl == "p2p":
                # peers will write into this rank's buffer b: whatever the caller enqueued so far (its reads of the frame
                # gathered into b two calls ago) comes first
                self.comm_stream.wait_event(ev_user)
                self._push_slots(b)
            else:
                dist.all_gather_into_tensor(g.view(-1), g[self.rank].view(-1), group=self.group)
            self._ev_comm[b].record(self.comm_stream)
        self._comm_used[b] = True
        self.gather = self.gathers[b]
        return self.gather

    def finish(self) -> None:
        """Make the current stream wait for everything generate_pipelined() queued."""
        torch = self.torch
        if hasattr(self, "comm_stream"):
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(self.comm_stream)
            cur.wait_stream(self.stream)

    def tile_view(self, global_tile: int, name: str):
        """Field `name` of any tile, from the gathered buffer: [N*N, comps]."""
        r, l = self.layout.owner(global_tile)
        b, e = self.layout.field_range(name)
        comps = dict(FIELDS)[name]
        n2 = self.layout.N * self.layout.N
        return self.gather[r, b:e].view(self.layout.tiles_per_rank, n2, comps)[l]

    def close(self) -> None:
        if getattr(self, "_peer_bases", None):
            self.torch.cuda.synchronize(self.device)
            self._close_peers()         # drop the IPC mappings before the owners free their buffers
        if hasattr(self, "ocean"):
            self.ocean.close()
