"""torchrun --nproc-per-node W tools/tiles_check.py -- one process per GPU through the mw_tiles_* C ABI.

Every rank checks, for the peer-memory gather AND the ncclAllGather arm:
  * every gathered tile against a SINGLE-HANDLE mw.Ocean run of that tile's parameters (seed 1000 + tile, wind rotated
    45 deg * first tile of the owning rank) -- bit-equal (same kernels, same inputs);
  * blocking frames, pipelined frames read one frame late (the documented consumer pattern), and back-to-back frames
    with no host synchronisation in between (the bench's loop).
Prints one JSON line and TILES_CHECK OK / FAIL from rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import mistral_water_b200 as mw
from mistral_water_b200.tiles import FIELDS, ShardedTiles, tile_wind

N, TPR = int(os.environ.get("MW_CHECK_N", "256")), int(os.environ.get("MW_CHECK_TPR", "2"))
TIMES = [0.0, 1.7, 3.25, 60.0, 0.5, 2.0]


def single_handle_frames():
    """expected[t][global tile] = dict(field -> tensor), from plain single-tile-set handles on this device"""
    exp = {}
    for r in range(world):
        o = mw.Ocean(N, seed=1000 + r * TPR, tiles=TPR, wind=tile_wind((5.0, 3.0), r * TPR), device=local, device_ptrs=True)
        o.init_spectrum()
        for t in TIMES:
            bufs = {k: torch.empty(TPR * N * N * c, device=dev) for k, c in FIELDS}
            o.generate(t, bufs)
            o.sync()
            exp.setdefault(t, {})[r] = torch.cat([bufs[k] for k, _ in FIELDS])   # = the slot layout
        o.close()
    return exp


exp = single_handle_frames()
res = {"rank": rank}
for arm in ("peer", "nccl"):
    st = ShardedTiles(N, rank, world, tiles_per_rank=TPR, device=dev, gather=arm)
    res[arm + "_impl"] = st.gather_impl
    ok_block = ok_pipe = ok_stream = True
    with torch.cuda.stream(st.stream):
        # blocking
        for t in TIMES[:2]:
            g = st.generate(t)
            st.sync()
            ok_block &= all(bool(torch.equal(g[r], exp[t][r])) for r in range(world))
        # pipelined, consumed one frame late: gen(k+1) is enqueued before frame k is read
        prev = None
        for k, t in enumerate(TIMES):
            g = st.generate_pipelined(t)
            if prev is not None:
                st.tileset.wait(1)
                pg, pt = prev
                got = pg.clone()            # ordered on the user stream
                st.stream.synchronize()
                ok_pipe &= all(bool(torch.equal(got[r], exp[pt][r])) for r in range(world))
            prev = (g, t)
        st.finish(); st.stream.synchronize()
        ok_pipe &= all(bool(torch.equal(prev[0][r], exp[prev[1]][r])) for r in range(world))
        # back to back, no host synchronisation, then the last two frames
        seq = [TIMES[k % len(TIMES)] for k in range(9)]
        bufs = [st.generate_pipelined(t) for t in seq]
        st.finish(); st.sync()
        ok_stream &= all(bool(torch.equal(bufs[-1][r], exp[seq[-1]][r])) for r in range(world))
        ok_stream &= all(bool(torch.equal(bufs[-2][r], exp[seq[-2]][r])) for r in range(world))
        # the two halves apart
        g = st.generate_local(TIMES[2]); st.all_gather(); st.finish(); st.sync()
        ok_stream &= all(bool(torch.equal(g[r], exp[TIMES[2]][r])) for r in range(world))
    res[arm] = {"blocking": ok_block, "pipelined": ok_pipe, "back_to_back": ok_stream}
    dist.barrier()
    st.close()
    dist.barrier()
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    print(json.dumps(allres), flush=True)
    good = all(all(r[a].values()) for r in allres for a in ("peer", "nccl")) and all(r["peer_impl"] == "peer" and r["nccl_impl"] == "nccl" for r in allres)
    print("TILES_CHECK", "OK" if good else "FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
