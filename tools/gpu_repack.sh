mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"repack|sgd_ema" -c 12 --csv --log-file gpurun_out/repack.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/repack.log 2>&1
grep -E "repack|sgd" gpurun_out/repack.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
