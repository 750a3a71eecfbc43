#!/usr/bin/env python
"""bench.py — headline benchmark of the Flux2Core denoising hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[1] — Klein 4B text-to-image 1024x1024 (4096 image + 512 text
tokens), 4 Euler steps bf16 + small-decoder VAE decode, random-init weights, synthetic inputs. One bench "step" = one
image = 4 DiT forwards + 4 Euler updates + unpack/BN-denorm/unpatchify + VAE decode + uint8 post-process.
N > 1 (torchrun): image-parallel — every rank generates its own images with replicated weights, no data-path
collective (SURVEY.md §8e mode 1), weak scaling; time = max over ranks.

`value`  : DiT steps/s, whole job, inputs resident in HBM (device pointers through the C ABI).
`e2e`    : the same through the C ABI with pinned HOST buffers (H2D of latents + text embeddings and D2H of the uint8
           image + final latents inside the timed region).
`roofline`: the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event time on the launch stream.
`cpu_baseline` / --impl reference: the restated reference CPU path (oracle, PyTorch-CPU fp32) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

# stdout carries exactly one JSON line. Libraries that write to file descriptor 1 on their own (NCCL prints "NCCL version ..." when
# the box exports NCCL_DEBUG=VERSION) are sent to stderr: fd 1 is re-pointed at fd 2 and the result line goes to a saved copy.
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: str) -> None:
    _RESULT_OUT.write(line + "\n")
    _RESULT_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "flux-2-swift-mlx_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

HEIGHT = WIDTH = 1024
S_TXT = 512
NUM_STEPS = 4
METRIC = "dit_steps_per_sec"
UNIT = "steps/s"


def peaks(rate_mult: int = 1):
    """Roofline denominators. rate_mult: tensor-pipe rate of the operand type relative to bf16 (fp8 x2, fp4 x4) — there is no
    measured fp8 / fp4 library number on this pool, so the block-scaled modes are normalised to the measured bf16 figure
    scaled by the nominal rate ratio (said so in peak_source)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    mult = "" if rate_mult == 1 else f" x {rate_mult} (nominal fp{8 if rate_mult == 2 else 4} : bf16 tensor rate)"
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops": rate_mult * d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)), "hbm": d.get("hbm_gbs", 6650.0),
                "src": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" + mult}
    return {"tflops": rate_mult * 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md sustained figure)" + mult}


def measured_traffic(kernel_class: str):
    """DRAM bytes per launch of a kernel class from the committed ncu pass (profiles/r01_traffic.json, written by
    tools/summarize_ncu.py traffic from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`), else None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(path)).get(kernel_class)
    except Exception:
        return None


def dit_flops(cfg, S_img, S_txt=512):
    """Algorithmic FLOPs of one DiT forward (BASELINE.md §3 accounting)."""
    D, Hm = cfg.inner_dim, cfg.mlp_hidden
    S = S_img + S_txt
    g = lambda M, N, K: 2 * M * N * K
    gemm = g(S_img, D, cfg.in_channels) + g(S_txt, D, cfg.joint_attention_dim) + g(S_img, cfg.out_channels, D)
    for s in (S_img, S_txt):
        gemm += cfg.num_layers * (4 * g(s, D, D) + g(s, 2 * Hm, D) + g(s, D, Hm))
    gemm += cfg.num_single_layers * (g(S, 3 * D + 2 * Hm, D) + g(S, D, D + Hm))
    attn = (cfg.num_layers + cfg.num_single_layers) * 4 * S * S * D
    return gemm, attn


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{gpu_index}.csv")

    def start(self):
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hot = sorted(sm)[len(sm) // 2:]  # samples under load dominate the upper half
            out = {"sm_mhz": statistics.median(hot), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(threads: int):
    """Bounded sample of the reference CPU path: 1 double-stream + 2 single-stream Klein-4B blocks at S = 4608
    (the 1024^2 token count), PyTorch-CPU fp32 (fp32 activations as in the reference); scaled to a full DiT step by
    block counts (5 double + 20 single). Embedders / final layer (<0.5 % of FLOPs) are not in the sample."""
    import torch
    from oracle import flux2_oracle as O
    torch.set_num_threads(threads)
    cfg = O.klein_4b()
    small = O.DiTConfig(num_layers=1, num_single_layers=2, num_attention_heads=cfg.num_attention_heads,
                        joint_attention_dim=cfg.joint_attention_dim, guidance_embeds=False)
    D, Hm = cfg.inner_dim, cfg.mlp_hidden
    W = {}
    for k, (o, i) in O.dit_weight_shapes(small).items():
        if k.startswith(("transformerBlocks", "singleTransformerBlocks", "doubleStream", "singleStream")):
            W[k] = torch.empty(o, i).uniform_(-1.0 / math.sqrt(i), 1.0 / math.sqrt(i))
    S_img = (HEIGHT // 16) * (WIDTH // 16)
    img, txt = torch.randn(1, S_img, D), torch.randn(1, S_TXT, D)
    temb = torch.randn(1, D)
    cos, sin = O.rope_embeddings(torch.cat([O.text_position_ids(S_TXT), O.image_position_ids(HEIGHT, WIDTH)]))
    im, tm = O.modulation(W, "doubleStreamModulationImg", temb, 2, D), O.modulation(W, "doubleStreamModulationTxt", temb, 2, D)
    sm = O.modulation(W, "singleStreamModulation", temb, 1, D)
    with torch.no_grad():
        t0 = time.perf_counter()
        txt2, img2, _ = O.double_block(W, 0, small, img, txt, im, tm, cos, sin)
        t_double = time.perf_counter() - t0
        x = torch.cat([txt2, img2], dim=1)
        t0 = time.perf_counter()
        for i in range(2):
            x, _ = O.single_block(W, i, small, x, sm, cos, sin)
        t_single = (time.perf_counter() - t0) / 2
    step_s = cfg.num_layers * t_double + cfg.num_single_layers * t_single
    gemm, attn = dit_flops(cfg, S_img)
    return {"step_s": step_s, "t_double": t_double, "t_single": t_single, "tflops": (gemm + attn) / step_s / 1e12,
            "sample": "1 double-stream + 2 single-stream Klein-4B blocks at S=4608 (1024x1024), PyTorch-CPU fp32 restatement "
                      "of the reference path (oracle); scaled to one DiT step by block counts (5 double + 20 single)"}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_sample(threads)
    vals = [cpu_sample(threads) for _ in range(max(1, min(args.steps, 3)))]
    best = min(vals, key=lambda v: v["step_s"])
    v = 1.0 / best["step_s"]
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": best["step_s"] * 1e3 * NUM_STEPS, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Klein 4B t2i 1024x1024 (4096 img + 512 txt tokens), DiT step; CPU arm = restated reference path "
                               "(MLX CPU backend not runnable here: no swift / mlx)", "l2": "n/a (host)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": best["sample"],
                         "cpu_tflops": best["tflops"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def make_weights_on_gpu(ctx, cfg, vcfg, device):
    import torch
    from oracle import flux2_oracle as O  # shapes / synthetic VAE weights only (input generation, not compute)
    g = torch.Generator(device=device).manual_seed(0)
    for k, (o, i) in O.dit_weight_shapes(cfg).items():
        b = 1.0 / math.sqrt(i)
        w = torch.empty(o, i, device=device, dtype=torch.float32).uniform_(-b, b, generator=g).to(torch.bfloat16)
        ctx.set_tensor(k, w)
        del w
    ctx.load_weights(O.random_vae_weights(vcfg, seed=1))


def run_sp(args, cfg, rank, local_rank, world, device, dist):
    """BASELINE.json configs[3]: one denoising step of a large joint sequence, Ulysses sequence-parallel over all ranks.
    Strong scaling: total work fixed, value = DiT steps/s of the whole job."""
    import torch
    import flux2b
    H = W = args.res
    S_img = (H // 16) * (W // 16)
    ctx = flux2b.Context(dit=cfg, device=local_rank, quant=flux2b.QUANT[args.quant],
                         options={"keep_raw_weights": 0, "sp_mode": args.sp_mode, "native_mx": args.native_mx, "mx_bn": args.mx_bn,
                                  "gemm_cta_group": args.cta_group})
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    from oracle import flux2_oracle as O
    g = torch.Generator(device=device).manual_seed(0)   # same seed on every rank: replicated weights
    n_lora = 0
    for k, (o, i) in O.dit_weight_shapes(cfg).items():
        b = 1.0 / math.sqrt(i)
        w = torch.empty(o, i, device=device, dtype=torch.float32).uniform_(-b, b, generator=g)
        if args.lora and (".attn." in k or ".ff" in k):
            # BASELINE.json configs[4]: a rank-16 LoRA (A, B ~ N(0, 0.02), scale 1) merged into every attention / FF linear at load
            # time (W += B · A before the packer runs; the on-device dequant -> add -> requant merge is covered by the tests)
            A = torch.randn(16, i, device=device, generator=g) * 0.02
            Bm = torch.randn(o, 16, device=device, generator=g) * 0.02
            w += Bm @ A
            n_lora += 1
        ctx.set_tensor(k, w.to(torch.bfloat16))
        del w
    ctx.finalize()
    torch.cuda.empty_cache()
    if world > 1:
        ids = [flux2b.sp_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, device=device)
        ctx.sp_init(ids[0], rank, world)
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(28, S_img)
    sig = sched.sigmas[:2]
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42)).to(device)
    enc = torch.randn(1, S_TXT, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43)).to(torch.bfloat16).to(device)
    guidance = 4.0 if cfg.guidance_embeds else None
    ref_lat = ref_ids = None
    S_ref = 0
    if args.refs > 0:
        # image-to-image conditioning: `refs` reference images at the output resolution, T coordinates 10, 20, 30 ...
        # (LatentUtils.swift:324-346), token order [output | refs] (Flux2Pipeline.swift:1703)
        S_ref = args.refs * S_img
        ref_lat = torch.randn(1, S_ref, 128, generator=torch.Generator().manual_seed(44)).to(device)
        ref_ids = O.reference_position_ids([H // 16] * args.refs, [W // 16] * args.refs).to(torch.int32).to(device)

    def step():
        x = lat.clone()
        ctx.denoise(x, enc, sig, H, W, guidance=guidance, ref_latents=ref_lat, ref_ids=ref_ids)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.prof_enable(True); ctx.prof_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    prof = {k: ctx.prof_get(i) for k, i in (("gemm", flux2b.PROF_GEMM), ("attn", flux2b.PROF_ATTN), ("elem", flux2b.PROF_ELEMWISE),
                                            ("gemv", flux2b.PROF_GEMV), ("comm", flux2b.PROF_COMM))}
    launches = ctx.launch_count()
    clocks = sampler.stop() if rank == 0 else {}
    barrier()
    if rank == 0:
        gemm_f, attn_f = dit_flops(cfg, S_img + S_ref)
        pk = peaks(rate_mult(args))
        gp = prof["gemm"]
        achieved = gp["flops"] / (gp["ms"] * 1e-3) / 1e12 if gp["ms"] > 0 else 0.0
        emit(json.dumps({
            "metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": dtype_name(args), "data": "synthetic",
            "config": {"workload": f"{args.model} one denoising step at {H}x{W} ({S_img} img + {S_ref} reference + {S_TXT} txt tokens), "
                                   f"{dtype_name(args)}{', rank-16 LoRA merged into ' + str(n_lora) + ' linears' if n_lora else ''}, Ulysses "
                                   f"sequence-parallel over {world} rank(s), transport mode {args.sp_mode}",
                       "l2": "inputs larger than L2 (weights stream from HBM every step)"},
            "tflops_total": (gemm_f + attn_f) * args.steps / (ms * 1e-3) / 1e12,
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_kernel", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "traffic": None, "peak_source": pk["src"]},
            "kernel_classes_rank0": {k: {"ms_per_step": p["ms"] / args.steps, "launches_per_step": p["launches"] / args.steps}
                                     for k, p in prof.items()},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def rate_mult(args) -> int:
    if not args.native_mx:
        return 1
    return {"mxfp8": 2, "mxfp4": 4, "nvfp4": 4}.get(args.quant, 1)


def dtype_name(args) -> str:
    """arithmetic type of the dominant kernel's operands"""
    if args.quant == "bf16":
        return "bf16"
    if args.native_mx:
        return f"{args.quant} (block-scaled tcgen05 MMA, weights and on-the-fly activations)"
    return f"bf16 x dequant({args.quant}) (W-only)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="flux2b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="klein4b")
    ap.add_argument("--quant", default="bf16", choices=["bf16", "qint8", "int4", "mxfp8", "mxfp4", "nvfp4"],
                    help="TransformerQuantization of the DiT linears (quantized on the fly by the bit-exact packer)")
    ap.add_argument("--native-mx", type=int, default=0,
                    help="1 = block linears on tcgen05 block-scaled MMA (mxfp8 / mxfp4 / nvfp4 weights consumed as packed, activations "
                         "quantised on the fly); 0 = W-only x · dequant(W)^T through the 16-bit GEMM (the reference's arithmetic)")
    ap.add_argument("--mx-bn", type=int, default=0, help="N tile of the block-scaled GEMM (0 = auto / 128 / 256)")
    ap.add_argument("--cta-group", type=int, default=0, help="GEMM CTA group (0 = auto: CTA pairs / 1 / 2)")
    ap.add_argument("--sp", action="store_true",
                    help="Ulysses sequence-parallel mode: all ranks cooperate on ONE denoising step (strong scaling); "
                         "use with --model dev --res 2048 (BASELINE.json configs[3])")
    ap.add_argument("--sp-mode", type=int, default=1, help="0 = NCCL all-to-all, 1 = peer-memory stores fused into the kernels")
    ap.add_argument("--res", type=int, default=1024, help="square resolution in pixels (--sp mode)")
    ap.add_argument("--refs", type=int, default=0, help="number of reference images (image-to-image conditioning tokens); without --sp also times the KV-cached loop")
    ap.add_argument("--lora", action="store_true", help="--sp mode: merge a rank-16 LoRA into every attention / FF linear at load time")
    ap.add_argument("--profile-one", action="store_true",
                    help="bracket ONE image with cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import flux2b
    from oracle import flux2_oracle as O  # configs + FLOP accounting inputs + cpu_baseline checker leg

    if flux2b.device_count() < 1:
        raise SystemExit("bench.py needs an sm_100 GPU: flux2b has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    cfg = {"klein4b": O.klein_4b, "klein9b": O.klein_9b, "dev": O.flux2_dev}[args.model]()
    vcfg = O.vae_small_decoder()
    if args.sp:
        run_sp(args, cfg, rank, local_rank, world, device, dist)
        return
    ctx = flux2b.Context(dit=cfg, vae=vcfg, device=local_rank, quant=flux2b.QUANT[args.quant],
                         options={"keep_raw_weights": 0, "native_mx": args.native_mx, "mx_bn": args.mx_bn,
                                  "gemm_cta_group": args.cta_group})
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    make_weights_on_gpu(ctx, cfg, vcfg, device)
    ctx.finalize()
    torch.cuda.empty_cache()

    S_img = (HEIGHT // 16) * (WIDTH // 16)
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(NUM_STEPS, S_img)
    gen = torch.Generator().manual_seed(42 + rank)
    lat_host = torch.randn(1, S_img, 128, generator=gen).pin_memory()
    enc_host = torch.randn(1, S_TXT, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43)).to(torch.bfloat16).pin_memory()
    guidance = 4.0 if cfg.guidance_embeds else None
    lat_dev0 = lat_host.to(device)
    enc_dev = enc_host.to(device)
    rgb_dev = torch.empty(HEIGHT, WIDTH, 3, dtype=torch.uint8, device=device)
    rgb_host = torch.empty(HEIGHT, WIDTH, 3, dtype=torch.uint8).pin_memory()

    def one_image_device():
        x = lat_dev0.clone()
        ctx.generate(x, enc_dev, sched.sigmas, HEIGHT, WIDTH, rgb_out=rgb_dev, guidance=guidance)

    x_host = torch.empty_like(lat_host).pin_memory()

    def one_image_host():
        x_host.copy_(lat_host)
        ctx.generate(x_host, enc_host, sched.sigmas, HEIGHT, WIDTH, rgb_out=rgb_host, guidance=guidance)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    if args.profile_one:
        one_image_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        one_image_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for _ in range(max(args.warmup, 3)):
        one_image_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The timed region proper: K steps, nothing but the product's own launches in the stream.
    ctx.prof_reset()
    ms = timed(one_image_device, args.steps)
    launches = ctx.launch_count()
    # The same K steps once more with a CUDA-event pair around every launch (on the context's stream): per-kernel-class durations for
    # `roofline` / `kernel_classes`. Kept out of the headline region because ~1700 extra event records per image cost ~2 %.
    ctx.prof_enable(True)
    ctx.prof_reset()
    ms_events = timed(one_image_device, args.steps)
    prof = {k: ctx.prof_get(i) for k, i in (("gemm", flux2b.PROF_GEMM), ("attn", flux2b.PROF_ATTN), ("elem", flux2b.PROF_ELEMWISE),
                                            ("conv", flux2b.PROF_CONV), ("gemv", flux2b.PROF_GEMV), ("groupnorm", flux2b.PROF_GROUPNORM))}
    ctx.prof_enable(False)
    ctx.prof_reset()
    # DiT-only timing (explains `value`): 4 forwards + Euler on device
    def dit_only():
        x = lat_dev0.clone()
        ctx.denoise(x, enc_dev, sched.sigmas, HEIGHT, WIDTH, guidance=guidance)
    dit_only()
    ms_dit = timed(dit_only, args.steps)
    i2i = None
    if args.refs > 0:
        # image-to-image with `refs` reference images: the standard loop re-processes the reference tokens
        # every step ([output | refs], Flux2Pipeline.swift:1696-1767); the klein-9b-kv loop extracts their K / V once (:1565-1644)
        S_ref = args.refs * S_img
        ref_lat = torch.randn(1, S_ref, 128, generator=torch.Generator().manual_seed(44)).to(device)
        ref_ids = O.reference_position_ids([HEIGHT // 16] * args.refs, [WIDTH // 16] * args.refs).to(torch.int32).to(device)
        res = {}
        for name, kv in (("standard", False), ("kv_cached", True)):
            def run():
                x = lat_dev0.clone()
                ctx.denoise(x, enc_dev, sched.sigmas, HEIGHT, WIDTH, guidance=guidance, ref_latents=ref_lat, ref_ids=ref_ids, kv_cache=kv)
            run()
            res[name] = timed(run, args.steps) / args.steps
        i2i = {"refs": args.refs, "S_ref": S_ref, "ms_per_image_standard": res["standard"], "ms_per_image_kv_cached": res["kv_cached"],
               "kv_speedup": res["standard"] / res["kv_cached"], "steps": NUM_STEPS}
    for _ in range(2):
        one_image_host()
    ms_e2e = timed(one_image_host, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    n_img = args.steps * world
    ms_per_image = ms / args.steps
    value = n_img * NUM_STEPS / (ms * 1e-3)
    e2e_value = n_img * NUM_STEPS / (ms_e2e * 1e-3)
    pk = peaks(rate_mult(args))
    g = prof["gemm"]
    achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    gemm_f, attn_f = dit_flops(cfg, S_img)
    total_kernel_ms = sum(p["ms"] for p in prof.values())
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_image, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_name(args), "data": "synthetic",
        "config": {"workload": f"{args.model} t2i {HEIGHT}x{WIDTH} ({S_img} img + {S_TXT} txt tokens), {NUM_STEPS} Euler steps {dtype_name(args)} + small-decoder "
                               "VAE decode (f16) per image; random-init weights; one image per bench step; image-parallel over ranks",
                   "l2": "inputs larger than L2 (7.8 GB of weights stream through the 126 MB L2 every forward)",
                   "images_per_rank": args.steps,
                   "kernel_timing": "roofline / kernel_classes come from a second pass over the same K steps with a CUDA-event pair around every launch"},
        "images_per_sec": n_img / (ms * 1e-3),
        "ms_per_step_with_kernel_events": ms_events / args.steps,
        "dit_only_steps_per_sec": n_img * NUM_STEPS / (ms_dit * 1e-3),
        "dit_tflops_per_gpu": (gemm_f + attn_f) * NUM_STEPS * args.steps / (ms_dit * 1e-3) / 1e12,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(lat_host.numel() * 4 + enc_host.numel() * 2),
                "d2h_bytes_per_step": int(HEIGHT * WIDTH * 3 + lat_host.numel() * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 GEMM, all DiT linears)", "achieved": achieved,
                     "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"] if pk["tflops"] else None,
                     "traffic": measured_traffic("gemm") if args.quant == "bf16" and args.model == "klein4b" else None,
                     "algorithmic_bytes_per_launch": g["bytes"] / g["launches"] if g["launches"] else None,
                     "peak_source": pk["src"], "launches": g["launches"],
                     "share_of_kernel_time": g["ms"] / total_kernel_ms if total_kernel_ms else None},
        "kernel_classes": {k: {"ms_per_image": p["ms"] / args.steps, "launches_per_image": p["launches"] / args.steps,
                               "tflops": (p["flops"] / (p["ms"] * 1e-3) / 1e12) if p["ms"] > 0 and p["flops"] else None,
                               "gbs": (p["bytes"] / (p["ms"] * 1e-3) / 1e9) if p["ms"] > 0 and p["bytes"] else None}
                           for k, p in prof.items()},
        "clocks": clocks,
    }
    if i2i is not None:
        out["i2i"] = i2i
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_sample(os.cpu_count() or 1)
        out["cpu_baseline"] = {"value": 1.0 / c["step_s"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": c["sample"], "cpu_tflops": c["tflops"]}
    emit(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
