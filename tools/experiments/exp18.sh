#!/bin/bash
# elimination experiments on attention variant 4 (timing only; results are wrong by construction)
for v in base EXP_SKIP_EXP EXP_SKIP_PV EXP_SKIP_LDS EXP_SKIP_S base; do
  if [ $v = base ]; then unset FLUX2B_LIB; else export FLUX2B_LIB=$PWD/flux-2-swift-mlx_b200/csrc/build/ab/libflux2b_$v.so; fi
  python tools/gpu_probe.py --run attn_v4_big 2>&1 | grep PROBE_RESULT | sed "s/^/$v /" | cut -c1-200
done
