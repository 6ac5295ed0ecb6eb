timeout 300 python -m pytest tests/test_renderer_gpu.py tests/test_gerstner_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do echo "MW_PDL=$v"; MW_PDL=$v timeout 200 python tools/bench_extra.py --only renderer 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print(d['config'], d['us_per_frame'])
    except Exception: pass
"; done
