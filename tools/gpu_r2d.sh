mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/r2d_pytest.log 2>&1; tail -15 gpurun_out/r2d_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; tail -2 gpurun_out/r2d_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench.log 2>&1; tail -1 gpurun_out/r2d_bench.log | cut -c1-3000
