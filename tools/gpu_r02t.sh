set -x
out=gpurun_out/r02t
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 2 -c 3 -o $out/prof_f32 python tools/ncu_driver.py cast_llnl 1e9 > $out/ncu.log 2>&1
tail -3 $out/ncu.log
ls -la $out
