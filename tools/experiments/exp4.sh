#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "quantized or in_kernel or forward_only or affine" > gpurun_out/pytest_wq.log 2>&1
echo "pytest wq rc=$?"; tail -n 3 gpurun_out/pytest_wq.log
for q in int4 qint8 nvfp4; do
  timeout 600 python bench.py --model klein9b --quant $q --wq-inkernel 1 --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra \
    > gpurun_out/bench_k9_${q}_ink1.json 2> gpurun_out/bench_k9_${q}_ink1.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_k9_${q}_ink1.json").read().strip().splitlines()[-1])
print("$q", {k: d[k] for k in ("value", "ms_per_step")}, d["kernel_classes"]["gemm"], d.get("mem_gb"))
PY
done
