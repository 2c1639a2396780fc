# ncu --set full captures of the dominant kernels (one launch each, after 2 warm-up launches) -> gpurun_out/*.ncu-rep
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
timeout 300 $NCU -k regex:conv_tc_fold -o gpurun_out/fold_c16_r02 python tools/run_conv_once.py 16 4 112 112 80 fold > gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_fold -o gpurun_out/fold_c32_r02 python tools/run_conv_once.py 32 4 56 56 40 fold >> gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_kernel -o gpurun_out/conv_c64_r02 python tools/run_conv_once.py 64 4 28 28 20 std >> gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_kernel -o gpurun_out/conv_c128_r02 python tools/run_conv_once.py 128 4 14 14 10 std >> gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_wgrad -o gpurun_out/wgrad_c16_r02 python tools/run_conv_once.py 16 4 112 112 80 wgrad >> gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_wgrad -o gpurun_out/wgrad_c32_r02 python tools/run_conv_once.py 32 4 56 56 40 wgrad >> gpurun_out/ncu_full.log 2>&1
timeout 300 $NCU -k regex:conv_tc_wgrad -o gpurun_out/wgrad_c256_r02 python tools/run_conv_once.py 256 4 7 7 5 wgrad >> gpurun_out/ncu_full.log 2>&1
# bandwidth kernels: every launch of one small elementwise/loss/norm exercise, selected metrics only
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:"mask_mix|label_mix|pseudo_label|sgd_ema|mix_loss|bn_apply|bn_stats|bn_bwd|bn_fused|head1|conv_first|repack|cc_" -c 400 --csv --log-file gpurun_out/bandwidth_kernels_r02.csv python bench.py --steps 2 --warmup 1 --profile >> gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
