#!/bin/bash
# multi-GPU round: sequence-parallel parity (both transports) + strong-scaling bench. usage: gpurun --gpus N -- 'bash tools/gpu_sp_round.sh N [model res]'
set -u
N=${1:-2}; MODEL=${2:-klein4b}; RES=${3:-1024}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in ${CHECK_MODES-0 1}; do
  SP_MODE=$mode timeout 600 $TR --master-port $((29600+mode)) tools/sp_check.py > gpurun_out/sp_check_n${N}_m${mode}.log 2>&1
  echo "sp_check N=$N mode=$mode rc=$?"; grep SP_CHECK gpurun_out/sp_check_n${N}_m${mode}.log | cut -c1-400; tail -n 3 gpurun_out/sp_check_n${N}_m${mode}.log | cut -c1-300
done
for mode in ${BENCH_MODES-0 1}; do
  timeout 900 $TR --master-port $((29610+mode)) bench.py --gpus $N --sp --sp-mode $mode --model $MODEL --res $RES --steps ${STEPS:-5} --warmup 3 \
    > gpurun_out/bench_sp_${MODEL}_${RES}_n${N}_m${mode}.json 2> gpurun_out/bench_sp_n${N}_m${mode}.err
  echo "bench sp N=$N mode=$mode rc=$?"; cut -c1-1200 gpurun_out/bench_sp_${MODEL}_${RES}_n${N}_m${mode}.json; tail -n 3 gpurun_out/bench_sp_n${N}_m${mode}.err | cut -c1-300
done
