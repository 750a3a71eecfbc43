#!/bin/bash
# 8-GPU round (BASELINE.json configs[2] and [4]). usage: gpurun --gpus 8 -- 'bash tools/gpu_n8_round.sh'
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# sequence-parallel parity with native mxfp8 linears (NCCL transport, overlapped exchanges, K-split out projection)
SP_QUANT=mxfp8 SP_NATIVE=1 SP_MODE=0 timeout 300 $TR --master-port 29601 tools/sp_check.py > gpurun_out/sp_check_n${N}_mxfp8_native.log 2>&1
echo "sp_check rc=$?"; grep SP_CHECK gpurun_out/sp_check_n${N}_mxfp8_native.log | head -n 2 | cut -c1-400
# configs[4]: Dev 32B mxfp8, 1024^2 output + 3 reference images + LoRA, sequence-parallel over all ranks
for nm in ${NATIVE_MODES-1}; do
  timeout 900 $TR --master-port 29612 bench.py --gpus $N --sp --sp-mode 0 --model dev --res 1024 --quant mxfp8 --native-mx $nm --refs 3 --lora \
    --steps 3 --warmup 3 > gpurun_out/bench_sp_dev_mxfp8_i2i3_lora_n${N}_native$nm.json 2> gpurun_out/bench_cfg4_$nm.err
  echo "configs[4] native=$nm rc=$?"; cut -c1-1500 gpurun_out/bench_sp_dev_mxfp8_i2i3_lora_n${N}_native$nm.json; tail -n 2 gpurun_out/bench_cfg4_$nm.err | cut -c1-300
done
# configs[2]: Klein 9B nvfp4 (block-scaled tcgen05 GEMMs), image-parallel over all ranks
timeout 900 $TR --master-port 29613 bench.py --gpus $N --model klein9b --quant nvfp4 --native-mx 1 --steps 3 --warmup 3 \
  > gpurun_out/bench_k9_nvfp4_native_n${N}.json 2> gpurun_out/bench_cfg2.err
echo "configs[2] rc=$?"; cut -c1-1500 gpurun_out/bench_k9_nvfp4_native_n${N}.json; tail -n 2 gpurun_out/bench_cfg2.err | cut -c1-300
