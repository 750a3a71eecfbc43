#!/bin/bash
mkdir -p gpurun_out
# f3: klein-9b-kv, standard vs KV-cached loop with 1 and 3 reference images
for r in 1 3; do
  timeout 600 python bench.py --model klein9b --refs $r --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra > gpurun_out/bench_k9_refs$r.json 2> gpurun_out/bench_k9_refs$r.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_k9_refs$r.json").read().strip().splitlines()[-1])
print("refs $r", d.get("i2i"), round(d["value"], 2))
PY
done
# a14: Dev 32B qint8, packed weights only: resident memory and speed
timeout 900 python bench.py --model dev --quant qint8 --steps 2 --warmup 3 --no-cpu-baseline --no-sp-extra > gpurun_out/bench_dev_qint8.json 2> gpurun_out/bench_dev_qint8.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_dev_qint8.json").read().strip().splitlines()[-1])
    print("dev qint8", round(d["value"], 2), round(d["ms_per_step"], 1), "mem_gb", round(d["mem_gb"], 1), d["kernel_classes"]["gemm"])
except Exception as e:
    print("no result", e); print(open("gpurun_out/bench_dev_qint8.err").read()[-1200:])
PY
