set -x
out=gpurun_out/r02f
mkdir -p $out
timeout 900 python tools/diag_groups.py > $out/groups.log 2>&1; cat $out/groups.log
timeout 900 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch|  w  |pathCB" > $out/diag.log; cat $out/diag.log | cut -c1-300
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_nomargin.so 2 > $out/probe_nomargin.log 2>&1; cat $out/probe_nomargin.log
timeout 2400 python -m pytest tests/test_gpu_retrace.py tests/test_gpu_fullsize_parity.py -m gpu -q -s --tb=short > $out/pytest_retrace.log 2>&1
grep -v "^$" $out/pytest_retrace.log | cut -c1-700 | tail -50
