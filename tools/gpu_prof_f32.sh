# usage: bash tools/gpu_prof_f32.sh <tag>   -> gpurun_out/<tag>/prof_f32.ncu-rep (one launch of the plain fused FP32 kernel, 1e9 rays)
set -x
out=gpurun_out/$1
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 1 -c 1 -o $out/prof_f32 python tools/ncu_driver.py cast_llnl 1e9 > $out/ncu.log 2>&1
tail -3 $out/ncu.log
