set -x
out=gpurun_out/r02B
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_retrace.py tests/test_gpu_f32.py tests/test_gpu_fullsize_parity.py tests/test_gpu_parity.py tests/test_c_abi.py -x -q 2>&1 | tail -5 | tee $out/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > $out/bench.json 2> $out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02B/bench.json'))
for k in ['value','ms_per_step','presampled','e2e_records','e2e_passed']:
    print(k, json.dumps(d.get(k))[:500])
PY
