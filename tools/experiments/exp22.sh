#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/experiments/exp22.sh N'
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -n 3 gpurun_out/bench_n$N.err
timeout 900 python -m pytest tests/test_gpu_sp.py -m gpu -q > gpurun_out/pytest_sp.log 2>&1; tail -n 3 gpurun_out/pytest_sp.log
