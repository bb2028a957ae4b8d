set -x
out=gpurun_out/r02b
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s --tb=short -k "corner or statistically" > $out/pytest.log 2>&1
grep -v "^$" $out/pytest.log | cut -c1-400 | tail -120
