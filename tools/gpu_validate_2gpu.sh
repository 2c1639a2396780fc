# 2-GPU data-parallel bench:  gpurun --gpus 2 --timeout 200 -- "bash tools/gpu_validate_2gpu.sh"
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "exit code $?"; grep '^{"metric' gpurun_out/bench_2gpu.log | cut -c1-200; tail -2 gpurun_out/bench_2gpu.log | cut -c1-200
