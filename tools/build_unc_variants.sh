# Development aid: libsart variants that keep a single decision group of the FP32 margin tests (SART_UNC_GROUPS), to
# measure how many rays each decision sends to the FP64 re-trace. Only kernels_f32.cu is recompiled.
set -e
cd "$(dirname "$0")/../solaraxionraytracing_b200"
python -m solaraxionraytracing_b200.build >/dev/null 2>&1 || (cd .. && python -m solaraxionraytracing_b200.build >/dev/null)
for g in 0 1 2 3 4 5 6 7 8 9; do
  (
  mask=$((1 << g))
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -ccbin /usr/bin/g++ -I ../include -DSART_UNC_GROUPS=${mask}u -c csrc/kernels_f32.cu -o build/k32_g$g.o 2>/dev/null
  objs=$(ls build/*.o | grep -v "kernels_f32.cu.o" | grep -v "k32_g")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -cudart static -o libsart_g$g.so $objs build/k32_g$g.o
  ) &
done
wait
ls -la libsart_g*.so | wc -l
