#!/bin/bash
# bash tools/push_call.sh W [cfg ...] -- the peer arm's push engines at W GPUs (gather alone and pipelined steps, tools/gather_probe.py),
# then the bit-for-bit tile check (tools/tiles_check.py).  A cfg is engine:ctas:chunk:priority (ce | sm:32 | tma:64:16384 | tma:16:16384:0) or "nccl";
# MW_PUSH_CHECK = the cfg the check runs with (default tma:5:49152 -- ragged last chunk, uneven deal).
set -u
N=${1:-4}; shift
CFGS=${*:-"tma:64:16384 tma:32:16384 sm:32 ce nccl"}
OUT=gpurun_out; mkdir -p $OUT
PORT=29511
setcfg() { IFS=: read -r e c k pr <<< "$1"; export MW_TILES_PUSH=$e; [ -n "${c:-}" ] && export MW_TILES_PUSH_CTAS=$c || unset MW_TILES_PUSH_CTAS; [ -n "${k:-}" ] && export MW_TILES_PUSH_CHUNK=$k || unset MW_TILES_PUSH_CHUNK; [ -n "${pr:-}" ] && export MW_TILES_PUSH_PRIO=$pr || unset MW_TILES_PUSH_PRIO; }
: > $OUT/push_probe_$N.jsonl; : > $OUT/push_probe.err
for cfg in $CFGS; do
  PORT=$((PORT+1))
  if [ "$cfg" = nccl ]; then export MW_PROBE_ARMS=nccl; unset MW_TILES_PUSH; else export MW_PROBE_ARMS=peer; setcfg $cfg; fi
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT tools/gather_probe.py 2>>$OUT/push_probe.err | grep '^{' >> $OUT/push_probe_$N.jsonl
done
for cfg in ${MW_PUSH_CHECK:-tma:5:49152}; do
  PORT=$((PORT+1)); setcfg $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT tools/tiles_check.py 2>&1 | tail -1
done
cat $OUT/push_probe_$N.jsonl; tail -5 $OUT/push_probe.err
