set -x
out=gpurun_out/r02H
mkdir -p $out
timeout 300 python tools/bench_configs.py 2 4 > $out/cfg4.log 2>&1; cat $out/cfg4.log
timeout 600 python -m pytest tests/test_gpu_retrace.py tests/test_gpu_fast.py tests/test_gpu_f32.py tests/test_gpu_parity.py -x -q -s -k "mass" 2>&1 | tail -14
