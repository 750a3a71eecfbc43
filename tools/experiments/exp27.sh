#!/bin/bash
# GEMM tile order (N fastest for the K = 9216 / 12288 out projections): correctness subset, then the bench with both orders
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "gemm or dit_forward or blocks_at_model_width or denoise or staged or native" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_sub.log
for o in auto 0; do
  if [ $o = auto ]; then unset FLUX2B_GEMM_ORDER; else export FLUX2B_GEMM_ORDER=$o; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-sp-extra --no-cpu-baseline > gpurun_out/bench_order_$o.json 2> gpurun_out/bench_order_$o.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_order_$o.json").read().strip().splitlines()[-1])
print("order=$o", round(d["value"],2), round(d["ms_per_step"],1), "gemm", round(d["kernel_classes"]["gemm"]["ms_per_image"],1), round(d["kernel_classes"]["gemm"]["tflops"]), "attn", round(d["kernel_classes"]["attn"]["ms_per_image"],1), d["clocks"]["sm_mhz"])
PY
done
