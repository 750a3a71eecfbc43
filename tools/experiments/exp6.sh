#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "conv2d or vae" > gpurun_out/pytest_conv.log 2>&1
echo "pytest conv rc=$?"; tail -n 8 gpurun_out/pytest_conv.log
timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_halo.log 2>&1; tail -n 9 gpurun_out/probe_conv_halo.log
bash tools/gpu_round.sh bench_nosp
