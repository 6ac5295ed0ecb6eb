#!/bin/bash
# bash tools/fallback_check.sh W -- bench.py at W GPUs as the driver launches it, once normally and once with the peer arm forced to fail at
# connect time (MW_TILES_FAIL_PEER_CONNECT): the line must then carry the ncclAllGather arm's value and say so.
set -u
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/fallback_$2.json 2> $OUT/fallback_$2.err; echo "$2 rc=$?"; }
run 29581 normal
MW_TILES_FAIL_PEER_CONNECT=1 run 29582 forced
python - <<'PY'
import json
for k in ("normal", "forced"):
    try:
        d = json.load(open(f"gpurun_out/fallback_{k}.json")); m = d["multi_gpu"]
        print(k, round(d["value"] / 1e9, 2), round(d["ms_per_step"], 4), m["default_arm"], {a: (v.get("ms_per_step") or v.get("unavailable")) for a, v in m["arms"].items()}, m.get("default_arm_fallback", ""), round(d["e2e"]["value"] / 1e9, 2))
    except Exception as e:
        print(k, "FAILED", repr(e)); print(open(f"gpurun_out/fallback_{k}.err").read()[-1500:])
PY
