#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 0 1; do
  if [ $f = 1 ]; then export FLUX2B_GEMM_FAKE_HALF_B=1; else unset FLUX2B_GEMM_FAKE_HALF_B; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-sp-extra --no-cpu-baseline > gpurun_out/bench_fake$f.json 2> gpurun_out/bench_fake$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_fake$f.json").read().strip().splitlines()[-1])
print("fake=$f", round(d["value"],2), round(d["ms_per_step"],1), "gemm", round(d["kernel_classes"]["gemm"]["ms_per_image"],1), round(d["kernel_classes"]["gemm"]["tflops"]), "attn", round(d["kernel_classes"]["attn"]["ms_per_image"],1), d["clocks"]["sm_mhz"])
PY
done
