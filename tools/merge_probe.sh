#!/bin/bash
# single GPU: IA3_FIT_MERGE (small rounds run entirely as team tasks) against latency, the 20-step run and steady state
for m in 0 32 74 148; do
  for steps in 20 96; do
    IA3_FIT_MERGE=$m timeout 200 python bench.py --steps $steps --warmup 5 --no-full-check --no-pageable --no-cpu > gpurun_out/merge_${m}_$steps.log 2>&1
    python - $m $steps <<'PY'
import json, sys
m, steps = sys.argv[1:3]
try:
    j = json.loads(open(f"gpurun_out/merge_{m}_{steps}.log").read().strip().splitlines()[-1])
    fa = j['config']['fit_stage_one_stack_alone']
    print(f"merge {m:4s} steps {steps:3s}: value {j['value']/1e3:6.0f}k {j['ms_per_step']:6.2f} ms | e2e {j['e2e']['value']/1e3:6.0f}k {j['e2e']['ms_per_step']:6.2f} ms | alone {j['config']['latency_ms_one_stack_alone']:.1f} ms, fit {fa['device_ms']:.1f} ms in {fa['engine']['rounds']} rounds, team tasks {fa['engine']['team_tasks']}")
except Exception as e:
    print(m, steps, "failed", e)
PY
  done
done
