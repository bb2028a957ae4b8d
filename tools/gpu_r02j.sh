set -x
out=gpurun_out/r02j
mkdir -p $out
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch" > $out/diag.log; cat $out/diag.log | cut -c1-300
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_nomargin.so 2 > $out/probe_nomargin.log 2>&1; cat $out/probe_nomargin.log
timeout 2400 python -m pytest tests/test_gpu_retrace.py tests/test_c_abi.py tests/test_gpu_fullsize_parity.py -m gpu -q --tb=short > $out/pytest.log 2>&1
grep -v "^$" $out/pytest.log | cut -c1-500 | tail -40
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -5 $out/bench.err; cat $out/bench.json
