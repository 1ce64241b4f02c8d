#!/bin/bash
# usage: tools/scale_probe.sh N   -- runs bench.py on N GPUs under a few host-side settings (same box), prints value / e2e / ms
N=$1
run() {  # tag, env..., -- extra args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-full-check --no-pageable "$@" > gpurun_out/probe_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/probe_{tag}.log").read().strip().splitlines()[-1])
    print(f"{tag:24s} value {j['value']/1e3:8.0f}k  {j['ms_per_step']:6.2f} ms/step | e2e {j['e2e']['value']/1e3:8.0f}k {j['e2e']['ms_per_step']:6.2f} ms (floor {j['e2e']['pcie_floor_ms']:.2f}) | host cpu {j['config']['host_cpu_ms_per_step']:.2f} ms")
except Exception as e:
    print(tag, "failed", e)
PY
}
nproc
export IA3_BENCH_VERBOSE=1
if [ "$2" = "short" ]; then
run unpinned IA3_BENCH_PIN=0 --
run pinned IA3_BENCH_PIN=1 --
run pinned_switch IA3_BENCH_PIN=1 IA3_SWITCH_INTERVAL=0.0005 --
run pinned_inflight16 IA3_BENCH_PIN=1 -- --inflight 16
run pinned_chunk8 IA3_BENCH_PIN=1 IA3_FIT_CHUNK=8 --
run pinned_s64 IA3_BENCH_PIN=1 -- --steps 64
grep -h "^rank" gpurun_out/probe_pinned.log | sort | head -8
exit 0
fi
run unpinned IA3_BENCH_PIN=0 --
run pinned IA3_BENCH_PIN=1 --
run pinned_inflight12 IA3_BENCH_PIN=1 -- --inflight 12
run pinned_switch IA3_BENCH_PIN=1 IA3_SWITCH_INTERVAL=0.0005 --
run pinned_s64 IA3_BENCH_PIN=1 -- --steps 64
N=1
run single X=1 --
