mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_conv_shapes.py -k "stride2 or s2 or down or up or first or same_direct" -q --timeout 300 > gpurun_out/s2_pytest.log 2>&1; tail -4 gpurun_out/s2_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines 2>&1 | tail -1 | cut -c1-330
