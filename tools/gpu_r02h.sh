set -x
out=gpurun_out/r02h
mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q --tb=short > $out/pytest_all.log 2>&1
grep -v "^$" $out/pytest_all.log | cut -c1-500 | tail -60
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch" > $out/diag.log; cat $out/diag.log | cut -c1-300
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_nomargin.so 2 > $out/probe_nomargin.log 2>&1; cat $out/probe_nomargin.log
