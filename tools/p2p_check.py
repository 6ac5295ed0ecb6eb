"""torchrun --nproc-per-node 2 tools/p2p_check.py  -- the peer-memory all-gather against ncclAllGather and against
single handles: every rank must end with every tile's fields, bit-identical to the NCCL result, blocking and pipelined."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from mistral_water_b200.tiles import ShardedTiles

N, TPR = 256, 3
res = {"rank": rank}
os.environ["MW_GATHER"] = "p2p"
a = ShardedTiles(N, rank, world, tiles_per_rank=TPR, device=dev)
os.environ["MW_GATHER"] = "nccl"
b = ShardedTiles(N, rank, world, tiles_per_rank=TPR, device=dev)
res["impl"] = [a.gather_impl, b.gather_impl]
res["p2p_error"] = a.p2p_error
ok = True
# blocking
for t in (0.0, 1.7):
    ga = a.generate(t).clone(); gb = b.generate(t).clone()
    torch.cuda.synchronize(); dist.barrier()
    ok &= bool(torch.equal(ga, gb)) and bool(ga.abs().sum() > 0)
    # every slot differs from every other (different seeds / winds)
    ok &= all(not torch.equal(ga[i], ga[j]) for i in range(world) for j in range(i))
res["blocking_equal"] = ok
# pipelined: frames 0..5, check each frame's buffer after finish()
ok2 = True
outs_a, outs_b = [], []
for st, outs in ((a, outs_a), (b, outs_b)):
    for k in range(6):
        g = st.generate_pipelined(0.25 * k)
        if k >= 1:
            pass
        st.finish()
        torch.cuda.synchronize()
        outs.append(g.clone())
    dist.barrier()
for x, y in zip(outs_a, outs_b):
    ok2 &= bool(torch.equal(x, y))
res["pipelined_equal"] = ok2
# back-to-back pipelined frames without intermediate finish (what bench.py does), then the last two frames
for st in (a, b):
    for k in range(8):
        st.generate_pipelined(0.1 * k)
    st.finish(); torch.cuda.synchronize(); dist.barrier()
res["stream_equal"] = bool(torch.equal(a.gathers[1], b.gathers[1])) and bool(torch.equal(a.gathers[0], b.gathers[0]))
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    print(json.dumps(allres), flush=True)
    good = all(r["blocking_equal"] and r["pipelined_equal"] and r["stream_equal"] for r in allres)
    print("P2P_CHECK", "OK" if good else "FAIL", flush=True)
a.close(); b.close()
dist.barrier()
dist.destroy_process_group()
