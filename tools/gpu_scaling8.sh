# 8-GPU confirmation of the current build: bench.py at N = 8 and BASELINE config 5 at full size (the 1/2/4-GPU points are
# in tools/gpu_scaling.sh; this one keeps the charged box time short).
set -x
out=gpurun_out/${1:-scale8}
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus $ngpu --steps 5 --warmup 3 > $out/bench_n$ngpu.json 2> $out/bench_n$ngpu.err
tail -2 $out/bench_n$ngpu.err; cat $out/bench_n$ngpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29511 tools/run_config5.py --rays ${2:-1e11} --check 1e9 --out $out > $out/config5_n$ngpu.json 2> $out/config5.err
tail -3 $out/config5.err; cat $out/config5_n$ngpu.json
