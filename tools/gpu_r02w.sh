set -x
out=gpurun_out/r02w
mkdir -p $out
SART_F32_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32x2 --launch-skip 1 -c 1 -o $out/prof_f32x2 python tools/ncu_driver.py cast_llnl 1e9 > $out/ncu.log 2>&1
tail -3 $out/ncu.log
ls -la $out
