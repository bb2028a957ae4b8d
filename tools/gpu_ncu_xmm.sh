# Full ncu capture of the compacting f32 kernel on BabyIAXO+XMM (1e9 rays), via tools/perf_probe.py.
set -x
out=gpurun_out/${1:-r01i}
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mc_f32_compact --launch-skip 1 -c 1 -o $out/prof_f32_xmm python tools/perf_probe.py 2 > $out/ncu_xmm.log 2>&1
tail -3 $out/ncu_xmm.log
