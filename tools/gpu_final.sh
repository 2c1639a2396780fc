# final validation of a build: full GPU suite, the three workload benches, cold + warm ncu launch lists of the LA step
mkdir -p gpurun_out
T=${1:-r02final}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=12 > gpurun_out/${T}_pytest.log 2>&1; tail -22 gpurun_out/${T}_pytest.log | cut -c1-160
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
for WL in la acdc pancreas; do
timeout 600 python bench.py --workload $WL --steps 30 --warmup 5 --no-baselines > gpurun_out/${T}_bench_$WL.log 2>&1; tail -1 gpurun_out/${T}_bench_$WL.log | cut -c1-260
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_ncu_list.log 2>&1; tail -1 gpurun_out/${T}_ncu_list.log | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/launches_${T}_warm.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/${T}_ncu_warm.log 2>&1; tail -1 gpurun_out/${T}_ncu_warm.log | cut -c1-120
