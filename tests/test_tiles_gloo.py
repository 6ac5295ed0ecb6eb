"""world_size-2 CPU (gloo) test of the multi-GPU path's host logic: tile -> rank sharding, slot layout
and the single in-place all-gather.  The CUDA producer is replaced by a stub that fills a rank's slot
with a pattern derived from the parameters ShardedTiles hands it (seed, wind, tiles) -- the collective,
offsets and views are the product code."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, tpr, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mistral_water_b200.tiles import FIELDS, ShardedTiles

        def make_stub(rp):
            def run(t, views):
                for name, comps in FIELDS:
                    v = views[name].view(rp["tiles"], N * N, comps)
                    for l in range(rp["tiles"]):
                        # value encodes (seed of the tile, field, t)
                        v[l] = float(rp["seed"] + l) + 0.001 * comps + t
            return run

        st = ShardedTiles(N, rank, world, tiles_per_rank=tpr, base_seed=1000, device=torch.device("cpu"),
                          make_generator=make_stub)
        assert st.rank_params["seed"] == 1000 + rank * tpr
        g = st.generate(0.5)
        ok = True
        for gt in range(world * tpr):
            for name, comps in FIELDS:
                tv = st.tile_view(gt, name)
                ok &= tv.shape == (N * N, comps)
                ok &= bool(torch.all(tv == float(1000 + gt) + 0.001 * comps + 0.5))
        ok &= g.shape == (world, tpr * N * N * 7)
        q.put((rank, ok, st.rank_params["wind"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_allgather_layout():
    world, N, tpr = 2, 32, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, tpr, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _ in res)
    # rank 1's first tile is global tile 2 -> wind rotated by 90 degrees
    assert np.allclose(res[1][2], (-3.0, 5.0))
