i=0
for cfg in "IA3_FIT_GRID_LATE=2368 IA3_FIT_GRID_TEAM=296 IA3_FIT_GRID_TEAM_LATE=296 IA3_FIT_CHUNK=8" "IA3_FIT_CHUNK=6" "IA3_FIT_GRID_LATE=148 IA3_FIT_GRID_TEAM_LATE=37 IA3_FIT_CHUNK=4" "IA3_FIT_GRID_LATE=592 IA3_FIT_GRID_TEAM_LATE=74 IA3_FIT_CHUNK=3"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 128 --warmup 5 --no-cpu --no-pageable --inflight 48 > gpurun_out/r2i_bench_$i.log 2> gpurun_out/r2i_bench_$i.err
  python -c "
import json,sys
l=json.loads(open('gpurun_out/r2i_bench_$i.log').read().strip().splitlines()[-1])
print('$cfg:', round(l['value']), round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value']), round(l['e2e']['ms_per_step'],2), 'alone', round(l['config']['latency_ms_one_stack_alone'],1), 'cpu', round(l['config']['host_cpu_ms_per_step'],2), 'launches/step', l['gpu_launches']//128)
"
done
