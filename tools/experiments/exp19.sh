#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "attention or dit_forward or kv or denoise" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_attn.log
for rep in 1 2; do for which in new old; do
  if [ $which = old ]; then export FLUX2B_LIB=$PWD/flux-2-swift-mlx_b200/csrc/build/ab/libflux2b_attn_old.so; else unset FLUX2B_LIB; fi
  for pr in attn_v4_big attn_v4_dev16k attn_v4_tail; do
    python tools/gpu_probe.py --run $pr 2>&1 | grep PROBE_RESULT | sed "s/^/$which $rep /" | cut -c1-170
  done
done; done
