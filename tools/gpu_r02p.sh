set -x
out=gpurun_out/r02p
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_retrace.py tests/test_gpu_xray_source.py tests/test_gpu_f32.py -m gpu -q --tb=short -s -k "generic or xray or f32" > $out/pytest.log 2>&1
grep -v "^$" $out/pytest.log | cut -c1-600 | tail -30
