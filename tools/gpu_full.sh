# full validation: GPU tests, smoke, bench (one JSON line), ncu launch list of one eager step.  usage: bash tools/gpu_full.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-baselines > gpurun_out/${TAG}_bench.log 2>&1; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${TAG}_ncu_list.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_list.log | cut -c1-200
