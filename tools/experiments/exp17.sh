#!/bin/bash
mkdir -p gpurun_out
for mb in 64 128 256 1024; do
  timeout 600 python bench.py --model klein9b --quant int4 --wq-inkernel 2 --wq-stage-mb $mb --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra > gpurun_out/bench_k9_int4_mb$mb.json 2> gpurun_out/bench_k9_int4_mb$mb.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_k9_int4_mb$mb.json").read().strip().splitlines()[-1])
    print("int4 staged chunk $mb MB", {k: round(d[k],2) for k in ("value", "ms_per_step")}, round(d["kernel_classes"]["gemm"]["ms_per_image"],1), d["kernel_classes"]["gemm"]["launches_per_image"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("no result", e); print(open("gpurun_out/bench_k9_int4_mb$mb.err").read()[-1500:])
PY
done
