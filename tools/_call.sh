set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -5
MW_PDL=2 timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "full_sizes or golden or tile_group" 2>&1 | tail -3
for v in "MW_PDL=0" "MW_PDL=1" "MW_PDL=2" "MW_PDL=0 MW_ROWS_MINB=4" "MW_PDL=1 MW_ROWS_MINB=4"; do
  env $v timeout 120 python tools/pdl_sweep.py 2>&1 | tail -1 | tee -a gpurun_out/pdl_sweep.jsonl
done
