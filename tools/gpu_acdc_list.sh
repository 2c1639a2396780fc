mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/launches_acdc_warm.csv python bench.py --workload acdc --steps 2 --warmup 1 --profile > gpurun_out/acdc_ncu_warm.log 2>&1; tail -1 gpurun_out/acdc_ncu_warm.log | cut -c1-120
