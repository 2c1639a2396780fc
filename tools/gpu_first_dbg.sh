mkdir -p gpurun_out
for c in "fwd 2d" "fwd 3d" "wgrad 2d" "wgrad 3d"; do
  echo "== $c"; timeout 100 python tools/debug_first.py $c 2>&1 | tail -2
done
bash tools/gpu_first.sh
