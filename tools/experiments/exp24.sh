#!/bin/bash
# setmaxnreg (dec 88 / inc 208) on attention variant 4 with S prefetch, live-in values consumed before the reallocation (variant library e)
export FLUX2B_LIB=$PWD/flux-2-swift-mlx_b200/csrc/build/ab/libflux2b_attn_e.so
for pr in attn_v4_small attn_v4_big attn_v4_dev16k; do
  timeout 120 python tools/gpu_probe.py --run $pr 2>&1 | grep -E "PROBE_RESULT|timeout tag" | sort | uniq -c | sort -rn | head -2 | sed "s/^/e /" | cut -c1-200
done
unset FLUX2B_LIB
python tools/gpu_probe.py --run attn_v4_big 2>&1 | grep PROBE_RESULT | sed "s/^/base /" | cut -c1-200
