# Full ncu capture of the fused f32 kernel with the alias sampler (skips the two first launches of sart_create and one warm-up step).
set -x
out=gpurun_out/${1:-r01s}
mkdir -p $out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_short.json 2> $out/bench_short.err
tail -c 600 $out/bench_short.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 -o $out/prof_f32_alias python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-presampled --sampler alias > $out/ncu_alias.log 2>&1
tail -2 $out/ncu_alias.log
