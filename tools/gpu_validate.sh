set -x
mkdir -p gpurun_out/r01b
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01b/pytest_gpu.log
cat gpurun_out/r01b/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r01b/bench.json 2> gpurun_out/r01b/bench.err
rc=$?
tail -3 gpurun_out/r01b/bench.err; cat gpurun_out/r01b/bench.json
if [ $rc -eq 0 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01b/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_fast -c 1 -o gpurun_out/r01b/prof_fast python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-presampled > gpurun_out/r01b/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_presampled -c 1 -o gpurun_out/r01b/prof_presampled python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01b/ncu_full2.log 2>&1
fi
ls -la gpurun_out/r01b
