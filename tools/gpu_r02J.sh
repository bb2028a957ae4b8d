set -x
out=gpurun_out/r02J
mkdir -p $out
rm -f $out/probe.log
for v in libsart.so libsart_rad12.so libsart.so libsart_rad12.so; do
  echo "== $v" >> $out/probe.log
  timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/$v 2 >> $out/probe.log 2>&1
done
cat $out/probe.log
