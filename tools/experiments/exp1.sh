#!/bin/bash
mkdir -p gpurun_out
for bo in 0 1 2; do
  FLUX2B_HALO_BO=$bo timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k conv2d > gpurun_out/conv_bo$bo.log 2>&1
  echo "BO=$bo: $(tail -n 1 gpurun_out/conv_bo$bo.log)"
done
timeout 900 python tools/gpu_probe.py attn_v3_big attn_v4 attn_v3_dev16k > gpurun_out/probe_attn4.log 2>&1; tail -n 12 gpurun_out/probe_attn4.log
timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_halo.log 2>&1; tail -n 9 gpurun_out/probe_conv_halo.log
FLUX2B_CONV_HALO=0 timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_pertap.log 2>&1; tail -n 9 gpurun_out/probe_conv_pertap.log
