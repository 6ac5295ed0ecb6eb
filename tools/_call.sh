for g in p2p nccl; do
  MW_GATHER=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/bench8_$g.err | tee gpurun_out/bench8_$g.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$g', d['value'], d['ms_per_step'], d['multi_gpu'])"
  tail -3 gpurun_out/bench8_$g.err
done
