mkdir -p gpurun_out
(timeout 300 python tools/debug_conv_tc.py > gpurun_out/dbg_fwd.log 2>&1; timeout 300 python tools/debug_conv_tc.py --wgrad > gpurun_out/dbg_wg.log 2>&1; timeout 300 python tools/debug_conv_tc.py --s2 > gpurun_out/dbg_s2.log 2>&1)
grep -h "WORST\|FAIL\|Error\|error" gpurun_out/dbg_*.log | head -20
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_graph.log 2>&1; tail -2 gpurun_out/bench_graph.log
