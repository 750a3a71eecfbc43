#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_round.sh tests bench_nosp
FLUX2B_GEMM_TIMELINE=1 timeout 300 python tools/gpu_probe.py --run vaeconv_1024_96 > gpurun_out/conv_timeline_halo.log 2>&1
FLUX2B_CONV_HALO=0 FLUX2B_GEMM_TIMELINE=1 timeout 300 python tools/gpu_probe.py --run vaeconv_1024_96 > gpurun_out/conv_timeline_pertap.log 2>&1
grep -c timeline gpurun_out/conv_timeline_halo.log
