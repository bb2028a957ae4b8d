set -x
out=gpurun_out/r02m4
mkdir -p $out
rm -f $out/cfg4.log
for v in libsart.so libsart_m640.so libsart_m512.so; do
  echo "== $v" >> $out/cfg4.log
  SART_LIB_VARIANT=$PWD/solaraxionraytracing_b200/$v timeout 300 python tools/bench_configs.py 2 4 >> $out/cfg4.log 2>&1
done
cat $out/cfg4.log
