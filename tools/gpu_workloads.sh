mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log | cut -c1-200
for WL in la acdc pancreas; do
timeout 600 python bench.py --workload $WL --steps 30 --warmup 5 --no-baselines > gpurun_out/r2h_bench_$WL.log 2>&1; tail -1 gpurun_out/r2h_bench_$WL.log | cut -c1-260
done
