# 8 GPUs of one box: bench.py at N = 8 and N = 4 under torchrun (sart_allreduce per step), the reference arm, the C driver's
# one-process 8-handle all-reduce. Usage: gpurun --gpus 8 -- bash tools/gpu_scaling8.sh <tag>
set -x
out=gpurun_out/${1:-scale8}
mkdir -p $out
nvidia-smi -L | wc -l
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_$n.json 2> $out/bench_$n.err
tail -2 $out/bench_$n.err; cut -c1-1500 $out/bench_$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $out/reference_8.json 2> $out/reference_8.err; cut -c1-700 $out/reference_8.json
timeout 300 python -m pytest tests/test_c_abi.py -m gpu -q -s 2>&1 | grep -E "abi_driver|passed|failed" | tail -8
