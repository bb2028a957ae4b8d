set -x
out=gpurun_out/r02d
mkdir -p $out
timeout 900 python tools/diag_retrace.py cast_llnl babyiaxo_xmm > $out/diag.log 2>&1; cat $out/diag.log | cut -c1-400
