set -x
out=gpurun_out/r02l
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_image --launch-skip 2 -c 1 -o $out/prof_exact python tools/ncu_driver_exact.py cast_llnl > $out/ncu.log 2>&1
tail -3 $out/ncu.log
