# developer script: bench.py under a list of environment settings ("VAR=val VAR2=val2" per line on stdin)
while read -r envs; do
  echo "ENV [$envs]"
  env $envs python bench.py --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), round(d['ms_per_step']*1e3,1), d['roofline']['avg_launch_ms'], d['roofline']['pipeline']['frac'])"
done
