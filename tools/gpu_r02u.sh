set -x
out=gpurun_out/r02u
mkdir -p $out
rm -f $out/probe.log
for p in 0 1 0 1; do
  echo "== SART_F32_PAIR=$p" >> $out/probe.log
  SART_F32_PAIR=$p timeout 300 python tools/perf_probe.py 2 >> $out/probe.log 2>&1
done
cat $out/probe.log
