#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "gemv or modulation or timestep or dit_forward or in_kernel or staged or denoise" > gpurun_out/pytest_sub.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_sub.log
bash tools/gpu_round.sh bench_nosp
