#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "ln_modulate or dit_forward or native or blocks_at_model_width" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_sub.log
bash tools/gpu_round.sh ncu bench_nosp
