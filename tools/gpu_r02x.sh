set -x
out=gpurun_out/r02x
mkdir -p $out
rm -f $out/probe.log
for v in libsart.so libsart_x2_640.so; do
  echo "== $v SART_F32_PAIR=1" >> $out/probe.log
  SART_F32_PAIR=1 timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/$v 2 >> $out/probe.log 2>&1
done
cat $out/probe.log
