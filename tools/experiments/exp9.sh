#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 1 -c 1 -o gpurun_out/prof_attn4 -f python tools/prof_kernels.py attn4 > gpurun_out/ncu_attn4.log 2>&1
tail -n 2 gpurun_out/ncu_attn4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/prof_conv_halo2 -f python tools/gpu_probe.py --run vaeconv_1024_96 > gpurun_out/ncu_conv_halo2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/prof_conv_out3 -f python tools/gpu_probe.py --run vaeconv_1024_96_3 > gpurun_out/ncu_conv_out3.log 2>&1
ls -la gpurun_out/*.ncu-rep
bash tools/gpu_round.sh ncu
