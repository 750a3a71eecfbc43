#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conv2d or vae or gemv or modulation or dit_forward or timestep or denoise" > gpurun_out/pytest_sub.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_sub.log
timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_wres.log 2>&1; tail -n 8 gpurun_out/probe_conv_wres.log
FLUX2B_CONV_WRES=0 timeout 600 python tools/gpu_probe.py vaeconv_1024 > gpurun_out/probe_conv_nowres.log 2>&1; tail -n 4 gpurun_out/probe_conv_nowres.log
bash tools/gpu_round.sh bench_nosp
