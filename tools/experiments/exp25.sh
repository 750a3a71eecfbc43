#!/bin/bash
# last check of the final tree: GPU suite, smoke, the default bench line twice (box-to-box / run-to-run spread), reference arm
mkdir -p gpurun_out
bash tools/gpu_round.sh tests smoke
for i in 1 2; do
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final$i.json 2> gpurun_out/bench_final$i.err; echo "bench $i rc=$?"
done
bash tools/gpu_round.sh benchref
