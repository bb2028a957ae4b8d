# Full validation on one B200: GPU test-suite, smoke, bench line, reference arm, ncu launch list of the same bench command, and
# full ncu captures of the dominant fused kernel, of the HBM-bound pre-sampled kernel and of the exact kernel.
# Usage: bash tools/gpu_validate.sh <tag>
set -x
out=gpurun_out/${1:-validate}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/pytest_gpu.log
cat $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
rc=$?
tail -3 $out/bench.err; cat $out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; cat $out/bench_reference.json
if [ $rc -eq 0 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > $out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 2 -c 1 -o $out/prof_f32 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-presampled --no-configs > $out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_presampled_f32 -c 1 -o $out/prof_presampled_f32 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > $out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_image --launch-skip 1 -c 1 -o $out/prof_exact python tools/ncu_driver_exact.py cast_llnl > $out/ncu_full3.log 2>&1
fi
ls -la $out
