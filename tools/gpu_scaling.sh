# Scaling run on one box: bench.py at N = 1, 2, 4, 8 (as many as are visible) and BASELINE config 5 at full size.
set -x
out=gpurun_out/${1:-scale}
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -le $ngpu ] || continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-presampled > $out/bench_n1.json 2> $out/bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
  fi
  tail -2 $out/bench_n$n.err; cat $out/bench_n$n.json
done
rays=${2:-1e11}
if [ $ngpu -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29511 tools/run_config5.py --rays $rays --check 1e9 --out $out > $out/config5_n$ngpu.json 2> $out/config5.err
else
  timeout 900 python tools/run_config5.py --rays 1e10 --check 1e9 --out $out > $out/config5_n1.json 2> $out/config5.err
fi
tail -3 $out/config5.err; cat $out/config5_n$ngpu.json
ls -la $out
