#!/bin/bash
# single GPU: fit-engine knobs against the 20-step run and one stack alone.  usage: tools/knob_probe.sh "A=1 B=2" "C=3" ...
for kv in "$@"; do
  tag=$(echo "$kv" | tr ' =' '__')
  env $kv timeout 200 python bench.py --steps ${STEPS:-20} --warmup 5 --no-full-check --no-pageable --no-cpu > gpurun_out/knob_$tag.log 2>&1
  python - "$tag" "$kv" <<'PY'
import json, sys
tag, kv = sys.argv[1:3]
try:
    j = json.loads(open(f"gpurun_out/knob_{tag}.log").read().strip().splitlines()[-1])
    fa = j['config']['fit_stage_one_stack_alone']
    print(f"{kv:44s}: value {j['value']/1e3:6.0f}k {j['ms_per_step']:6.2f} ms | e2e {j['e2e']['value']/1e3:6.0f}k {j['e2e']['ms_per_step']:6.2f} ms | alone {j['config']['latency_ms_one_stack_alone']:.1f} ms, fit {fa['device_ms']:.1f} ms in {fa['engine']['rounds']} rounds")
except Exception as e:
    print(kv, "failed", e)
PY
done
