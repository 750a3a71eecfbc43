#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_round.sh tests
timeout 600 python tools/gpu_probe.py vaeconv attn_v3_big attn_v4_big attn_v4_dev16k gemm_big_cg2 gemm_out_cg2 gemm_m512_out > gpurun_out/probe_r2b.log 2>&1; tail -n 14 gpurun_out/probe_r2b.log
bash tools/gpu_round.sh bench_nosp
