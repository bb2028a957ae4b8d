# 2 GPUs: the C driver's sart_allreduce over 2 handles in one process, and bench.py under torchrun (sart_allreduce per step)
set -x
out=gpurun_out/r02n
mkdir -p $out
nvidia-smi -L
timeout 600 python -m pytest tests/test_c_abi.py -m gpu -q -s --tb=short 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench2.json 2> $out/bench2.err
tail -5 $out/bench2.err; cat $out/bench2.json | cut -c1-3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/ref2.json 2>> $out/bench2.err; cat $out/ref2.json | cut -c1-600
