set -x
out=gpurun_out/r02o
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_retrace.py tests/test_gpu_xray_source.py -m gpu -q --tb=short -x > $out/pytest.log 2>&1
grep -v "^$" $out/pytest.log | cut -c1-400 | tail -12
for v in "" solaraxionraytracing_b200/libsart_ex6.so solaraxionraytracing_b200/libsart_ex8.so; do timeout 300 python tools/perf_probe.py $v 0 2>&1 | tail -2; done
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=1 scale=1.0|retrace=0" | cut -c1-200
