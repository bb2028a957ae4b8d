# Full ncu capture of the timed fused kernel (skips the pilot launch sart_create makes for its autotune).
set -x
out=gpurun_out/${1:-r01c}
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_fast --launch-skip 2 -c 1 -o $out/prof_fast python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-presampled > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log
ls -la $out
