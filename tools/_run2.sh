mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread -k "cc or step or primit" > gpurun_out/pytest_gpu2.log 2>&1; tail -3 gpurun_out/pytest_gpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; tail -3 gpurun_out/bench_2gpu.log | cut -c1-1500
