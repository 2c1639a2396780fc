mkdir -p gpurun_out
timeout 120 tools/micro/norm_probe_bin | head -12
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-baselines > gpurun_out/r2e_bench.log 2>&1; tail -1 gpurun_out/r2e_bench.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/r2e_ncu_list.log 2>&1; tail -1 gpurun_out/r2e_ncu_list.log | cut -c1-200
