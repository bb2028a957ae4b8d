# Full ncu capture of the timed single-precision fused kernel (skips the pilot launch of sart_create).
set -x
out=gpurun_out/${1:-r01f}
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 -o $out/prof_f32 python bench.py --precision f32 --steps 1 --warmup 3 --no-cpu-baseline --no-presampled > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log
ls -la $out
