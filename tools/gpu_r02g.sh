set -x
out=gpurun_out/r02g
mkdir -p $out
timeout 300 python tools/diag_ray.py cast_llnl 6203998 > $out/ray.log 2>&1; cat $out/ray.log
M="smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_margin.csv python tools/ncu_driver.py cast_llnl > $out/ncu1.log 2>&1
SART_LIB=$PWD/solaraxionraytracing_b200/libsart_nomargin.so timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 --csv --log-file $out/ncu_nomargin.csv python tools/ncu_driver.py cast_llnl > $out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32 --launch-skip 3 -c 1 -o $out/prof_f32_margin python tools/ncu_driver.py cast_llnl > $out/ncu3.log 2>&1
ls -la $out
