# Round 2, first GPU call: the new full-size parity tests + baseline throughput of the day + atomics / replica experiments.
set -x
out=gpurun_out/r02a
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s 2>&1 | tail -60 > $out/pytest_fullsize.log
cat $out/pytest_fullsize.log
timeout 300 python tools/perf_probe.py 2 > $out/probe_base.log 2>&1; cat $out/probe_base.log
timeout 300 python tools/perf_probe.py solaraxionraytracing_b200/libsart_noatom.so 2 > $out/probe_noatom.log 2>&1; cat $out/probe_noatom.log
SART_IMG_REP_SKEW=2080 timeout 300 python tools/perf_probe.py 2 > $out/probe_skew2080.log 2>&1; cat $out/probe_skew2080.log
SART_IMG_REP_SKEW=520 SART_IMG_REPLICAS=16 timeout 300 python tools/perf_probe.py 2 > $out/probe_skew520_r16.log 2>&1; cat $out/probe_skew520_r16.log
SART_IMG_REPLICAS=1 timeout 300 python tools/perf_probe.py 2 > $out/probe_rep1.log 2>&1; cat $out/probe_rep1.log
