mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_steps_ckpt.py tests/test_gpu_dropin.py -k "mix_loss or ckpt or dropin or step or loss" -q --timeout 600 --durations=8 > gpurun_out/loss_pytest.log 2>&1; tail -14 gpurun_out/loss_pytest.log | cut -c1-200
timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines 2>&1 | tail -1 | cut -c1-330
