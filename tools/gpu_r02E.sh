set -x
out=gpurun_out/r02E
mkdir -p $out
timeout 300 python tools/perf_probe.py 0 2 > $out/probe.log 2>&1
cut -c1-330 $out/probe.log
timeout 1500 python -m pytest tests/test_gpu_retrace.py -x -q 2>&1 | tail -5 | tee $out/tests.log
