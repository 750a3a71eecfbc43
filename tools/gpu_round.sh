#!/bin/bash
# One gpurun call: GPU test-suite, bench (both arms), ncu launch list of one image. Outputs land in gpurun_out/.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tests|bench|ncu|full ...]'
set -u
mkdir -p gpurun_out
WHAT="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
for w in $WHAT; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
      tail -n 60 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
      echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json ;;
    ncu)
      timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv \
        --log-file gpurun_out/launches.csv python bench.py --profile-one > gpurun_out/ncu_launches.log 2>&1
      echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv ;;
    te_ncu)
      # launch list of ONE text-encoder prefill (Klein 4B's Qwen3-4B, 27 layers, 512 tokens)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
        --log-file gpurun_out/te_launches.csv python tools/te_bench.py qwen3_4b --profile-one > gpurun_out/ncu_te.log 2>&1
      echo "ncu te launches rc=$?"; wc -l gpurun_out/te_launches.csv ;;
    traffic)
      # DRAM bytes of EVERY launch of one image (cheap pass: two counters, no --set full)
      timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
        -c 2000 --csv --log-file gpurun_out/traffic.csv python bench.py --profile-one > gpurun_out/ncu_traffic.log 2>&1
      echo "ncu traffic rc=$?"; wc -l gpurun_out/traffic.csv ;;
    bench_k9)
      # BASELINE.json configs[2] on one GPU: Klein 9B nvfp4, W-only and native block-scaled
      for nm in 0 1; do
        timeout 900 python bench.py --model klein9b --quant nvfp4 --native-mx $nm --steps 3 --warmup 3 --no-cpu-baseline \
          > gpurun_out/bench_k9_nvfp4_native$nm.json 2> gpurun_out/bench_k9_$nm.err
        echo "bench k9 native=$nm rc=$?"; tail -c 2500 gpurun_out/bench_k9_nvfp4_native$nm.json; tail -n 3 gpurun_out/bench_k9_$nm.err
      done ;;
    sanitize)
      # compute-sanitizer over a small grouped forward + VAE decode (tools/sanitize_target.py): memcheck, racecheck, synccheck
      for tool in memcheck racecheck synccheck; do
        timeout 900 compute-sanitizer --tool $tool --print-limit 400 python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
        echo "sanitize $tool rc=$?"; grep -E "SANITIZE_OK|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | head -5
      done ;;
    bench_nosp)
      timeout 900 python bench.py --steps 5 --warmup 3 --no-sp-extra > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    bench_wq)
      # W-only quantized Klein 9B (BASELINE.json configs[2] weights, reference arithmetic): in-kernel dequantisation vs dense copies
      for q in int4 qint8 nvfp4; do for ink in 1 0; do
        timeout 600 python bench.py --model klein9b --quant $q --wq-inkernel $ink --steps 3 --warmup 3 --no-cpu-baseline --no-sp-extra \
          > gpurun_out/bench_k9_${q}_ink$ink.json 2> gpurun_out/bench_k9_${q}_ink$ink.err
        echo "bench k9 $q inkernel=$ink rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_k9_${q}_ink$ink.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["kernel_classes"]["gemm"], d.get("mem_gb"))
except Exception as e:
    print("no result:", e); print(open("gpurun_out/bench_k9_${q}_ink$ink.err").read()[-1500:])
PY
      done; done ;;
    tests_wq)
      timeout 900 python -m pytest tests -m gpu -q -k "quantized or in_kernel or forward_only or lora or prequantized or affine or conv2d or vae" > gpurun_out/pytest_wq.log 2>&1
      echo "pytest wq rc=$?"; tail -n 25 gpurun_out/pytest_wq.log ;;
    probe_r2)
      # round-2 kernel candidates against what they replace: attention variant 4 vs 3, halo-tile convolutions vs one box per tap
      timeout 900 python tools/gpu_probe.py attn_v3_big attn_v4 attn_v3_dev16k > gpurun_out/probe_attn4.log 2>&1; tail -n 12 gpurun_out/probe_attn4.log
      timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_halo.log 2>&1; tail -n 9 gpurun_out/probe_conv_halo.log
      FLUX2B_CONV_HALO=0 timeout 600 python tools/gpu_probe.py vaeconv > gpurun_out/probe_conv_pertap.log 2>&1; tail -n 9 gpurun_out/probe_conv_pertap.log ;;
    full)
      # top kernels, full sections (few launches each)
      timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'gemm_kernel|attn_kernel' -s 40 -c 8 -o gpurun_out/prof_full -f python bench.py --profile-one > gpurun_out/ncu_full.log 2>&1
      echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep ;;
    prof_kernels)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn_kernel' -s 2 -c 20 \
        -o gpurun_out/prof_kernels -f python tools/prof_kernels.py ${PROF_ARGS:-gemm_cg1 gemm_cg2 attn3} > gpurun_out/prof_kernels.log 2>&1
      echo "prof_kernels rc=$?"; tail -n 3 gpurun_out/prof_kernels.log ;;
    probe_attn)
      timeout 900 python tools/gpu_probe.py attn_v > gpurun_out/probe_attn.log 2>&1; tail -n 20 gpurun_out/probe_attn.log ;;
    tests_attn)
      timeout 900 python -m pytest tests -m gpu -q -k "attention or dit_forward or denoise or kv" > gpurun_out/pytest_attn.log 2>&1; tail -n 15 gpurun_out/pytest_attn.log ;;
    probe)
      timeout 1500 python tools/gpu_probe.py gemm_big gemm_ffin gemm_out attn_v1_big attn_v2_big conv3_big > gpurun_out/probe.log 2>&1; tail -n 20 gpurun_out/probe.log ;;
  esac
done
