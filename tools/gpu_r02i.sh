set -x
out=gpurun_out/r02i
mkdir -p $out
M="smsp__inst_executed.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_margin.csv python tools/ncu_driver.py cast_llnl > $out/ncu1.log 2>&1
SART_LIB=$PWD/solaraxionraytracing_b200/libsart_nomargin.so timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_nomargin.csv python tools/ncu_driver.py cast_llnl > $out/ncu2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_margin_xmm.csv python tools/ncu_driver.py babyiaxo_xmm > $out/ncu3.log 2>&1
SART_LIB=$PWD/solaraxionraytracing_b200/libsart_nomargin.so timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_nomargin_xmm.csv python tools/ncu_driver.py babyiaxo_xmm > $out/ncu4.log 2>&1
tail -n 4 $out/ncu_*.csv | cut -c1-50,180-400
timeout 600 python tools/diag_retrace.py cast_llnl babyiaxo_xmm 2>&1 | grep -E "retrace=|mismatch" > $out/diag.log; cat $out/diag.log | cut -c1-300
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_nomargin.so 2 > $out/probe_nomargin.log 2>&1; cat $out/probe_nomargin.log
timeout 2400 python -m pytest tests -m gpu -q --tb=short > $out/pytest_all.log 2>&1
grep -v "^$" $out/pytest_all.log | cut -c1-500 | tail -60
