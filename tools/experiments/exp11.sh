#!/bin/bash
# A/B on one box: attention variant 4, masked-body split (lib B) vs if-converted masking (product lib)
mkdir -p gpurun_out
AB=flux-2-swift-mlx_b200/csrc/build/ab/libflux2b_attn_split.so
for which in base split; do
  if [ $which = split ]; then export FLUX2B_LIB=$PWD/$AB; else unset FLUX2B_LIB; fi
  for pr in attn_v4_big attn_v4_big_p4 attn_v4_big_p2 attn_v4_big_p0; do
    python tools/gpu_probe.py --run $pr 2>&1 | grep PROBE_RESULT | sed "s/^/$which /" | cut -c1-160
  done
done
