set -x
out=gpurun_out/r02y
mkdir -p $out
timeout 300 python tools/perf_probe.py 2 > $out/probe.log 2>&1
cat $out/probe.log
timeout 1500 python -m pytest tests/test_gpu_retrace.py tests/test_gpu_f32.py tests/test_gpu_hole_effarea.py -x -q 2>&1 | tail -15 | tee $out/tests.log
