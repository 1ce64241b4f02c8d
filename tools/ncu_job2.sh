set -x
cd $GRAFT_REPO_ROOT 2>/dev/null || true
python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --inflight 1 > gpurun_out/r02b_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-pageable --inflight 1 > gpurun_out/r02b_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/corr_ncu.py > gpurun_out/r02b_plain_corr.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_spline|k_warp|k_mix|k_hot_count" -c 12 -o gpurun_out/r02b_corr python tools/corr_ncu.py > gpurun_out/r02b_ncu_corr.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r02b_*
