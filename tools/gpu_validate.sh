# GPU-box validation used in round 1:  gpurun --timeout 1000 -- "bash tools/gpu_validate.sh"  (logs land in gpurun_out/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
if ! grep -q " passed" gpurun_out/pytest_gpu.log || grep -q "failed" gpurun_out/pytest_gpu.log; then
  BCP_FUSED_STATS=0 timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu_nofuse.log 2>&1; echo NOFUSE; tail -3 gpurun_out/pytest_gpu_nofuse.log | cut -c1-300
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_graph.log 2>&1; tail -1 gpurun_out/bench_graph.log | cut -c1-250
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200

