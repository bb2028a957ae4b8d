set -x
out=gpurun_out/r02k
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_image --launch-skip 1 -c 1 -o $out/prof_exact python tools/ncu_driver_exact.py cast_llnl > $out/ncu.log 2>&1
tail -3 $out/ncu.log
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch" > $out/diag.log; cat $out/diag.log | cut -c1-300
