# Round 2: margin-triggered FP64 re-trace. Tests + throughput with / without margins.
set -x
out=gpurun_out/r02c
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_retrace.py -m gpu -q -s --tb=short -x > $out/pytest_retrace.log 2>&1
grep -v "^$" $out/pytest_retrace.log | cut -c1-600 | tail -60
timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s --tb=short > $out/pytest_fullsize.log 2>&1
grep -v "^$" $out/pytest_fullsize.log | cut -c1-600 | tail -40
timeout 300 python tools/perf_probe.py 2 > $out/probe_margins.log 2>&1; cat $out/probe_margins.log
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_nomargin.so 2 > $out/probe_nomargin.log 2>&1; cat $out/probe_nomargin.log
