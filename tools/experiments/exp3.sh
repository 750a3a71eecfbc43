#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "quantized or in_kernel or forward_only or lora or prequantized or affine" > gpurun_out/pytest_wq.log 2>&1
echo "pytest wq rc=$?"; tail -n 5 gpurun_out/pytest_wq.log
for q in int4 qint8 nvfp4; do
  timeout 600 python bench.py --model klein9b --quant $q --wq-inkernel 1 --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra \
    > gpurun_out/bench_k9_${q}_ink1.json 2> gpurun_out/bench_k9_${q}_ink1.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_k9_${q}_ink1.json").read().strip().splitlines()[-1])
print("$q", {k: d[k] for k in ("value", "ms_per_step")}, d["kernel_classes"]["gemm"], d.get("mem_gb"))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/prof_conv_halo -f python tools/gpu_probe.py --run vaeconv_1024_96 > gpurun_out/ncu_conv_halo.log 2>&1
FLUX2B_CONV_HALO=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -o gpurun_out/prof_conv_pertap -f python tools/gpu_probe.py --run vaeconv_1024_96 > gpurun_out/ncu_conv_pertap.log 2>&1
ls -la gpurun_out/*.ncu-rep
