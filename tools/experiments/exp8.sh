#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "quantized or in_kernel or staged or forward_only or affine or lora or prequantized" > gpurun_out/pytest_wq.log 2>&1
echo "pytest wq rc=$?"; tail -n 6 gpurun_out/pytest_wq.log
for q in int4 qint8; do for ink in 2; do
  timeout 600 python bench.py --model klein9b --quant $q --wq-inkernel $ink --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra \
    > gpurun_out/bench_k9_${q}_ink$ink.json 2> gpurun_out/bench_k9_${q}_ink$ink.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_k9_${q}_ink$ink.json").read().strip().splitlines()[-1])
    print("$q ink=$ink", {k: d[k] for k in ("value", "ms_per_step")}, d["kernel_classes"]["gemm"], d.get("mem_gb"), d["clocks"])
except Exception as e:
    print("no result", e); print(open("gpurun_out/bench_k9_${q}_ink$ink.err").read()[-1500:])
PY
done; done
