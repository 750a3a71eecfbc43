#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/prof_wq_int4 -f python tools/prof_kernels.py wq_int4 > gpurun_out/ncu_wq.log 2>&1
tail -n 3 gpurun_out/ncu_wq.log
FLUX2B_GEMM_TIMELINE=1 timeout 300 python tools/prof_kernels.py wq_int4 > gpurun_out/wq_timeline.log 2>&1
grep -c timeline gpurun_out/wq_timeline.log
