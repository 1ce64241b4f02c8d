set -x
cd $GRAFT_REPO_ROOT 2>/dev/null || true
python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --inflight 1 > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --inflight 1 > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/engine_trace.py c2 v4 > gpurun_out/r02_plain_trace.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fit_round -c 8 -o gpurun_out/r02_fit_round python tools/engine_trace.py c2 v4 > gpurun_out/r02_ncu_trace.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r02_*
