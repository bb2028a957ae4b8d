set -x
out=gpurun_out/r02e
mkdir -p $out
./tools/micro/ffma2 > $out/ffma2.log 2>&1; cat $out/ffma2.log
timeout 900 python tools/diag_groups.py > $out/groups.log 2>&1; cat $out/groups.log
timeout 2400 python -m pytest tests/test_gpu_retrace.py -m gpu -q -s --tb=short > $out/pytest_retrace.log 2>&1
grep -v "^$" $out/pytest_retrace.log | cut -c1-700 | tail -70
