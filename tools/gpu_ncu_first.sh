mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
timeout 300 $NCU -k regex:conv_first_fwd_tma -o gpurun_out/first_fwd_r02 python tools/time_first.py > gpurun_out/ncu_first.log 2>&1
timeout 300 $NCU -k regex:conv_first_wgrad_tma -o gpurun_out/first_wgrad_r02 python tools/time_first.py >> gpurun_out/ncu_first.log 2>&1
tail -3 gpurun_out/ncu_first.log; ls -la gpurun_out/first_*.ncu-rep
