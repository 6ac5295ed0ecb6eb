for v in "" _w4; do for g in 1 16; do echo "VARIANT [$v] GROUP $g"; MW_LIB_SUFFIX=$v MW_GROUP_TILES=$g python bench.py --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), round(d['ms_per_step']*1e3,1), d['roofline']['avg_launch_ms'], d['roofline']['pipeline']['frac'])"; done; done
