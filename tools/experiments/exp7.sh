#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 2500 gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
timeout 900 python -m pytest tests/test_gpu_sp.py -m gpu -q > gpurun_out/pytest_sp.log 2>&1; tail -n 5 gpurun_out/pytest_sp.log
