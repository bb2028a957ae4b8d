set -x
out=gpurun_out/r02m
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_retrace.py tests/test_gpu_xray_source.py tests/test_gpu_angular_scan.py -m gpu -q --tb=short -x > $out/pytest.log 2>&1
grep -v "^$" $out/pytest.log | cut -c1-400 | tail -25
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch" > $out/diag.log; cat $out/diag.log | cut -c1-300
timeout 300 python tools/perf_probe.py 0 > $out/probe_exact.log 2>&1; cat $out/probe_exact.log
