#!/bin/bash
# Round-end evidence on ONE GPU (all outputs under gpurun_out/; summaries are copied into profiles/ by hand afterwards):
# full GPU test-suite, smoke, default bench (with the Ulysses extra at N = 1), reference arm, ncu launch list / DRAM traffic of every
# launch of one image / full sections of the top kernels, compute-sanitizer, Klein 9B nvfp4 W-only + native, text-encoder bench.
set -u
mkdir -p gpurun_out
bash tools/gpu_round.sh tests smoke
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
bash tools/gpu_round.sh benchref ncu traffic full sanitize
for nm in 0 1; do
  timeout 900 python bench.py --model klein9b --quant nvfp4 --native-mx $nm --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra \
    > gpurun_out/bench_k9_nvfp4_native$nm.json 2> gpurun_out/bench_k9_$nm.err
  echo "bench k9 nvfp4 native=$nm rc=$?"; tail -c 600 gpurun_out/bench_k9_nvfp4_native$nm.json
done
timeout 600 python tools/te_bench.py qwen3_4b > gpurun_out/te_bench.json 2> gpurun_out/te_bench.err; echo "te_bench rc=$?"; tail -c 600 gpurun_out/te_bench.json
