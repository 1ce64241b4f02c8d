#!/bin/bash
# single GPU: effect of the interpreter's thread switch interval on the many-stacks-in-flight pipeline
for si in 0.005 0.0005 0.0001 0.00002; do
  for steps in 20 96; do
    IA3_SWITCH_INTERVAL=$si timeout 200 python bench.py --steps $steps --warmup 5 --no-full-check --no-pageable > gpurun_out/switch_${si}_$steps.log 2>&1
    python - $si $steps <<'PY'
import json, sys
si, steps = sys.argv[1:3]
try:
    j = json.loads(open(f"gpurun_out/switch_{si}_{steps}.log").read().strip().splitlines()[-1])
    print(f"interval {si:8s} steps {steps:3s}: value {j['value']/1e3:6.0f}k {j['ms_per_step']:6.2f} ms | e2e {j['e2e']['value']/1e3:6.0f}k {j['e2e']['ms_per_step']:6.2f} ms | host cpu {j['config']['host_cpu_ms_per_step']:.2f}")
except Exception as e:
    print(si, steps, "failed", e)
PY
  done
done
