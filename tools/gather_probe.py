"""torchrun --nproc-per-node W tools/gather_probe.py -- the all-gather of the tile set alone (no frame kernels), both arms, under
the environment's knobs (MW_TILES_PUSH_LANES, CUDA_DEVICE_MAX_CONNECTIONS): one JSON line from rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from mistral_water_b200.tiles import ShardedTiles

N, T, K = int(os.environ.get("MW_PROBE_N", "2048")), int(os.environ.get("MW_PROBE_TILES", "1")), 40
res = {"world": world, "N": N, "tiles": T, "lanes": os.environ.get("MW_TILES_PUSH_LANES", ""), "maxconn": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", ""),
       "push": os.environ.get("MW_TILES_PUSH", "ce"), "push_ctas": os.environ.get("MW_TILES_PUSH_CTAS", ""), "push_chunk": os.environ.get("MW_TILES_PUSH_CHUNK", ""), "prio": os.environ.get("MW_TILES_PUSH_PRIO", "")}
for arm in os.environ.get("MW_PROBE_ARMS", "peer,nccl").split(","):
    st = ShardedTiles(N, rank, world, tiles_per_rank=T, device=dev, gather=arm)
    with torch.cuda.stream(st.stream):
        st.generate_local(0.0); st.finish()
        for _ in range(5): st.all_gather()
        st.finish(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st.stream)
        for _ in range(K): st.all_gather()
        st.finish(); e1.record(st.stream)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # pipelined steps (gather under the next frame's generation)
        for k in range(3): st.generate_pipelined(0.1 * k)
        st.finish(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record(st.stream)
        for k in range(3 * K): st.generate_pipelined(0.016 * k)
        st.finish(); e1.record(st.stream)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / (3 * K)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    slot = st.layout.slot_bytes
    res[arm] = {"gather_ms": round(ms, 4), "busbw_gbs": round(slot * (world - 1) / ms / 1e6, 1), "step_ms": round(float(t.item()), 4)}
    st.close(); dist.barrier()
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
