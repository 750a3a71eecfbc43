#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; tail -c 3500 gpurun_out/bench_n8.json; tail -n 3 gpurun_out/bench_n8.err
timeout 900 python -m pytest tests/test_gpu_sp.py -m gpu -q > gpurun_out/pytest_sp.log 2>&1; tail -n 3 gpurun_out/pytest_sp.log
