mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m multigpu -q --durations=5 > gpurun_out/dp2_pytest.log 2>&1; tail -15 gpurun_out/dp2_pytest.log | cut -c1-400
for OV in 0 1; do
BCP_DP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-baselines > gpurun_out/dp2_bench_ov$OV.log 2>&1; tail -1 gpurun_out/dp2_bench_ov$OV.log | cut -c1-330
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-baselines 2>&1 | tail -1 | cut -c1-330
