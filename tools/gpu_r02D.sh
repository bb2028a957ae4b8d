set -x
out=gpurun_out/r02D
mkdir -p $out
timeout 300 python tools/perf_probe.py 0 2 > $out/probe.log 2>&1
cat $out/probe.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrace.py tests/test_gpu_fullsize_parity.py tests/test_gpu_xray_source.py -x -q 2>&1 | tail -5 | tee $out/tests.log
