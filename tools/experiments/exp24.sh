#!/bin/bash
# liveness / speed of setmaxnreg configurations on attention variant 4 with S prefetch (variant libraries):
# c = dec 152 only, d = dec 120 / inc 192
for which in c d; do
  export FLUX2B_LIB=$PWD/flux-2-swift-mlx_b200/csrc/build/ab/libflux2b_attn_$which.so
  for pr in attn_v4_small attn_v4_big; do
    timeout 120 python tools/gpu_probe.py --run $pr 2>&1 | grep -E "PROBE_RESULT|timeout tag" | sort | uniq -c | sort -rn | head -2 | sed "s/^/$which /" | cut -c1-200
  done
done
