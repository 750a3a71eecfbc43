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


def csrc_sha() -> str:
    """sha256 over the kernel sources (csrc/*.cu, *.cuh, ctx.h): what a profile is valid for. The GPU box has no .git, so
    profiles are stamped with this instead of a commit id."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "flux-2-swift-mlx_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel_class: str):
    """DRAM bytes per launch of a kernel class from the ncu pass of THIS tree (tools/gpu_round.sh traffic ->
    tools/summarize_ncu.py traffic -> profiles/r02_traffic.json, stamped with csrc_sha). A file measured on other kernel
    sources is not served: (None, why)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(path))
    except Exception:
        return None, "no ncu traffic pass committed for this round"
    if d.get("csrc_sha") != csrc_sha():
        return None, f"profiles/r02_traffic.json was measured on kernel sources {d.get('csrc_sha')}, this tree is {csrc_sha()}"
    return d.get(kernel_class), f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, csrc_sha {d.get('csrc_sha')}"


def ncu_class_ms():
    """Per-class kernel durations of one image from the committed ncu launch list (profiles/r02_launches.json), served only while
    the kernel sources match: pure kernel time, without the ~2 - 5 us an event pair adds to every launch."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_launches.json")))
    except Exception:
        return None
    if d.get("csrc_sha") != csrc_sha():
        return None
    return {k: v["ms"] for k, v in d["classes"].items()}


def dit_flops(cfg, S_img, S_txt=512):
    """Algorithmic FLOPs of one DiT forward (BASELINE.md §3 accounting) — flux2b.configs.dit_flops."""
    from flux2b import configs
    return configs.dit_flops(cfg, S_img, S_txt)


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
    return {"step_s": step_s, "t_double": t_double, "t_single": t_single, "tflops": (gemm + attn) / step_s / 1e12, "flops": (gemm + attn, 0),
            "sample": "1 double-stream + 2 single-stream Klein-4B blocks at S=4608 (1024x1024), PyTorch-CPU fp32 restatement "
                      "of the reference path (oracle); scaled to one DiT step by block counts (5 double + 20 single)"}


def run_reference(args, rank):
    """--impl reference: the restated reference CPU path on all host threads. One "step" = one bounded sample (3 real Klein-4B
    blocks at S = 4608, scaled to a DiT step by block counts): W warm-up samples, then EXACTLY K timed samples; the line says
    "extrapolated": true. A real full step (25 blocks, 15.5 GB of fp32 weights) takes ~17 s on 16 cores, K of them do not fit the
    few-minute budget of a bench run."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_sample(threads)
    vals = [cpu_sample(threads) for _ in range(max(1, args.steps))]
    step_s = sum(v["step_s"] for v in vals) / len(vals)
    v = 1.0 / step_s
    gemm, attn = vals[0]["flops"]   # (total FLOPs of one DiT step, 0)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": step_s * 1e3 * NUM_STEPS, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": True,
        "config": {"workload": "Klein 4B t2i 1024x1024 (4096 img + 512 txt tokens), DiT step; CPU arm = restated reference path "
                               "(MLX CPU backend not runnable here: no swift / mlx); each timed step is a bounded sample (1 double + 2 "
                               "single blocks, really computed) scaled to 5 + 20 blocks; ms_per_step = 4 DiT steps, no VAE decode",
                   "l2": "n/a (host)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": vals[0]["sample"],
                         "cpu_tflops": (gemm + attn) / step_s / 1e12},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def load_synthetic_dit(ctx, cfg, device, lora=False, seed=0):
    """random-init DiT weights U(-1/sqrt(in), 1/sqrt(in)) (MLX Linear default) generated on the GPU, keys / shapes from the
    product's own manifest (flux2b.configs.dit_weight_manifest). Same seed on every rank: replicated weights."""
    import torch
    from flux2b import configs
    g = torch.Generator(device=device).manual_seed(seed)
    n_lora = 0
    for k, (o, i) in configs.dit_weight_manifest(cfg).items():
        b = 1.0 / math.sqrt(i)
        w = torch.empty(o, i, device=device, dtype=torch.float32).uniform_(-b, b, generator=g)
        if lora and (".attn." in k or ".ff" in k):
            # BASELINE.json configs[4]: a rank-16 LoRA (A, B ~ N(0, 0.02), scale 1) merged into every attention / FF linear at load
            # time (W += B · A before the packer runs; the on-device dequant -> add -> requant merge is covered by the tests)
            A = torch.randn(16, i, device=device, generator=g) * 0.02
            Bm = torch.randn(o, 16, device=device, generator=g) * 0.02
            w += Bm @ A
            n_lora += 1
        ctx.set_tensor(k, w.to(torch.bfloat16))
        del w
    return n_lora


def load_synthetic_vae(ctx, vcfg, device, seed=1):
    """synthetic VAE decoder weights: fan-in-scaled uniform convs / linears, GroupNorm gamma ~ 1, beta ~ 0, BN mean ~ 0, var ~ 1"""
    import torch
    from flux2b import configs
    g = torch.Generator(device=device).manual_seed(seed)
    for k, shape in configs.vae_weight_manifest(vcfg).items():
        if k.startswith("latentBatchNorm"):
            t = 0.1 * torch.randn(shape, device=device, generator=g) if k.endswith("Mean") else 1.0 + 0.1 * torch.rand(shape, device=device, generator=g)
        elif ".norm" in k or "groupNorm" in k or "convNormOut" in k:
            t = (1.0 if k.endswith(".weight") else 0.0) + 0.1 * torch.randn(shape, device=device, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if k.endswith(".bias"):   # bias bound follows the layer's fan-in: look at the weight's shape
                wshape = configs.vae_weight_manifest(vcfg)[k[:-len(".bias")] + ".weight"]
                fan_in = 1
                for d in wshape[1:]:
                    fan_in *= d
            bnd = 1.0 / math.sqrt(fan_in)
            t = torch.empty(shape, device=device, dtype=torch.float32).uniform_(-bnd, bnd, generator=g)
            if k.endswith(".weight"):
                t = t.half().float()
        ctx.set_tensor(k, t.contiguous())


def sp_measure(args, cfg, rank, local_rank, world, device, dist, res, refs, modes, steps, lora=False, model_name="dev"):
    """One denoising step of a large joint sequence, Ulysses sequence-parallel over all ranks of the job (BASELINE.json
    configs[3] / [4]); strong scaling: the work is fixed, value = DiT steps/s of the whole job. At world == 1 the same step runs
    on one GPU (the point scaling efficiency is computed against). Returns a dict per transport mode:
      value, ms_per_step (CUDA events, max over ranks), kernels_ms (rank 0: sum of the compute kernels' own durations from a second
      pass with an event pair per launch), exposed_comm_ms = ms_per_step - kernels_ms, parity_vs_single_gpu (rel-L2 of the sharded
      forward against the same context running the whole forward alone, option sp_disable)."""
    import torch
    import flux2b
    H = W = res
    S_img = (H // 16) * (W // 16)
    ctx = flux2b.Context(dit=cfg, device=local_rank, quant=flux2b.QUANT[args.quant],
                         options={"keep_raw_weights": 0, "native_mx": args.native_mx, "mx_bn": args.mx_bn, "gemm_cta_group": args.cta_group,
                                  "wq_inkernel": getattr(args, "wq_inkernel", 2)})
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    n_lora = load_synthetic_dit(ctx, cfg, device, lora=lora)
    ctx.finalize()
    torch.cuda.empty_cache()
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(28, S_img)
    sig = sched.sigmas[:2]
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42)).to(device)
    enc = torch.randn(1, S_TXT, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43)).to(torch.bfloat16).to(device)
    guidance = 4.0 if cfg.guidance_embeds else None
    ref_lat = ref_ids = None
    S_ref = 0
    if refs > 0:
        # image-to-image conditioning: `refs` reference images at the output resolution, T coordinates 10, 20, 30 ...
        # (LatentUtils.swift:324-346), token order [output | refs] (Flux2Pipeline.swift:1703)
        S_ref = refs * S_img
        ref_lat = torch.randn(1, S_ref, 128, generator=torch.Generator().manual_seed(44)).to(device)
        ref_ids = torch.from_numpy(flux2b.reference_position_ids([H // 16] * refs, [W // 16] * refs)).to(torch.int32).to(device)

    def step(x=None):
        x = lat.clone() if x is None else x
        ctx.denoise(x, enc, sig, H, W, guidance=guidance, ref_latents=ref_lat, ref_ids=ref_ids)
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gemm_f, attn_f = dit_flops(cfg, S_img + S_ref)
    out = {"workload": f"{model_name} one denoising step at {H}x{W} ({S_img} img + {S_ref} reference + {S_TXT} txt tokens), {dtype_name(args)}"
                       f"{', rank-16 LoRA merged into ' + str(n_lora) + ' linears' if n_lora else ''}, Ulysses sequence-parallel over {world} rank(s)",
           "scaling": "strong", "tflop_per_step": (gemm_f + attn_f) / 1e12, "steps": steps, "n_gpus": world, "modes": {}}
    single = None
    for mode in (modes if world > 1 else [None]):
        if world > 1:
            ctx.set_option("sp_mode", mode)
            ids = [flux2b.sp_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0, device=device)
            ctx.sp_init(ids[0], rank, world)
        for _ in range(3):
            step()
        ctx.prof_enable(False); ctx.prof_reset()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = ctx.launch_count()
        # second pass: an event pair around every launch -> the compute kernels' own time on rank 0
        ctx.prof_enable(True); ctx.prof_reset()
        barrier()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        prof = {k: ctx.prof_get(i) for k, i in (("gemm", flux2b.PROF_GEMM), ("attn", flux2b.PROF_ATTN), ("elem", flux2b.PROF_ELEMWISE),
                                                ("gemv", flux2b.PROF_GEMV), ("comm", flux2b.PROF_COMM))}
        ctx.prof_enable(False); ctx.prof_reset()
        kernels_ms = sum(prof[k]["ms"] for k in ("gemm", "attn", "elem", "gemv")) / steps
        parity = None
        if world > 1:
            # the sharded step against the same context running the whole step alone (each rank computes it redundantly)
            x_sp = step().float().cpu()
            ctx.set_option("sp_disable", 1)
            if single is None:
                single = step().float().cpu()
            ctx.set_option("sp_disable", 0)
            # relative to the UPDATE of the step (x - x0 = dt * velocity), not to the latents: one Euler step moves x by a few
            # percent, so a difference measured against |x| would be diluted by that factor
            x0 = lat.float().cpu().double()
            parity = float((x_sp.double() - single.double()).norm() / (single.double() - x0).norm())
        barrier()
        ms_step = ms / steps
        gp = prof["gemm"]
        out["modes"]["single_gpu" if mode is None else ("nccl_all_to_all" if mode == 0 else "peer_memory_fused")] = {
            "value": 1e3 / ms_step, "unit": UNIT, "ms_per_step": ms_step, "kernels_ms": kernels_ms,
            "exposed_comm_ms": max(0.0, ms_step - kernels_ms), "exposed_comm_frac": max(0.0, ms_step - kernels_ms) / ms_step,
            "parity_vs_single_gpu": parity, "tflops_total": (gemm_f + attn_f) / (ms_step * 1e-3) / 1e12,
            "gemm_tflops_rank0": gp["flops"] / (gp["ms"] * 1e-3) / 1e12 if gp["ms"] > 0 else None,
            "gpu_launches": int(launches),
            "kernel_classes_rank0": {k: {"ms_per_step": p["ms"] / steps, "launches_per_step": p["launches"] / steps} for k, p in prof.items()}}
    best = max(out["modes"].values(), key=lambda m: m["value"])
    out.update({"value": best["value"], "unit": UNIT, "ms_per_step": best["ms_per_step"], "kernels_ms": best["kernels_ms"],
                "exposed_comm_ms": best["exposed_comm_ms"], "parity_vs_single_gpu": best["parity_vs_single_gpu"]})
    ctx.close()
    del ctx
    torch.cuda.empty_cache()
    return out


def k9_measure(args, rank, local_rank, world, device, dist, steps=2):
    """BASELINE.json configs[2] in the same bench line (key `configs2`): Klein 9B with nvfp4 weights at 1024x1024, 4 Euler steps + VAE decode
    per image, image-parallel over all ranks — W-only (the reference's arithmetic: x . dequant(W)^T, packed weights only) and native
    block-scaled tcgen05 MMA (activations quantised on the fly), side by side with what each mode's parity is."""
    import torch
    import flux2b
    from flux2b import configs
    cfg, vcfg = configs.klein_9b(), configs.vae_small_decoder()
    S_img = (HEIGHT // 16) * (WIDTH // 16)
    sched = flux2b.FlowMatchEulerScheduler()
    sched.set_timesteps(NUM_STEPS, S_img)
    lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42 + rank)).to(device)
    enc = torch.randn(1, S_TXT, cfg.joint_attention_dim, generator=torch.Generator().manual_seed(43)).to(torch.bfloat16).to(device)
    rgb = torch.empty(HEIGHT, WIDTH, 3, dtype=torch.uint8, device=device)
    out = {"workload": f"klein9b nvfp4 t2i {HEIGHT}x{WIDTH} ({S_img} img + {S_TXT} txt tokens), {NUM_STEPS} Euler steps + small-decoder VAE decode per image; "
                       f"random-init weights; image-parallel over {world} rank(s)", "steps": steps, "n_gpus": world, "modes": {}}
    notes = {"w_only": "x . dequant(W)^T, the reference's arithmetic (per-block rel-L2 <= 2e-3 vs the oracle on the dequantized weights); packed weights are the only resident copy",
             "native_block_scaled": "tcgen05 block-scaled MMA on MLX's packed nvfp4 bytes, activations quantised on the fly to nvfp4: ~5.9e-2 rel-L2 from the reference arithmetic by construction (its own parity protocol: exact vs the emulation)"}
    for name, native in (("w_only", 0), ("native_block_scaled", 1)):
        free0, _ = torch.cuda.mem_get_info()
        ctx = flux2b.Context(dit=cfg, vae=vcfg, device=local_rank, quant=flux2b.QUANT["nvfp4"], options={"keep_raw_weights": 0, "native_mx": native})
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        load_synthetic_dit(ctx, cfg, device)
        load_synthetic_vae(ctx, vcfg, device)
        ctx.finalize()
        torch.cuda.empty_cache(); torch.cuda.synchronize()
        free1, _ = torch.cuda.mem_get_info()

        def one():
            x = lat.clone()
            ctx.generate(x, enc, sched.sigmas, HEIGHT, WIDTH, rgb_out=rgb)
        for _ in range(3):
            one()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        out["modes"][name] = {"value": steps * world * NUM_STEPS / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                              "images_per_sec": steps * world / (ms * 1e-3), "mem_gb": (free0 - free1) / 1e9,
                              "gpu_launches": int(ctx.launch_count()), "parity": notes[name]}
        ctx.close()
        del ctx
        torch.cuda.empty_cache()
    return out


def run_sp(args, cfg, rank, local_rank, world, device, dist):
    """--sp: the sequence-parallel measurement as the bench line itself (BASELINE.json configs[3] / [4])."""
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r = sp_measure(args, cfg, rank, local_rank, world, device, dist, args.res, args.refs, [args.sp_mode], max(args.steps, 1),
                   lora=args.lora, model_name=args.model)
    clocks = sampler.stop() if rank == 0 else {}
    if rank == 0:
        m = list(r["modes"].values())[0]
        pk = peaks(rate_mult(args))
        achieved = m["gemm_tflops_rank0"] or 0.0
        emit(json.dumps({
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps"], "warmup": 3,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype_name(args), "data": "synthetic",
            "config": {"workload": r["workload"] + f", transport mode {args.sp_mode}",
                       "l2": "inputs larger than L2 (weights stream from HBM every step)"},
            "tflops_total": m["tflops_total"], "gpu_launches": m["gpu_launches"], "sp": r,
            "roofline": {"bound": "tensor", "kernel": "gemm_kernel", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tflops"], "traffic": None, "peak_source": pk["src"]},
            "kernel_classes_rank0": m["kernel_classes_rank0"], "clocks": clocks}))
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
    return f"bf16 x dequant({args.quant}) (W-only, {('dense 16-bit copies', 'packed weights only, dequantized inside the kernels', 'packed weights only, dequantized per launch into a 16-bit stage')[getattr(args, 'wq_inkernel', 2)]})"


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
    ap.add_argument("--wq-inkernel", type=int, default=2,
                    help="W-only quantized layers: 2 = packed weights only, many-row GEMMs dequantize the layer into a 16-bit stage per launch; "
                         "1 = packed weights only, always dequantized inside the GEMM / GEMV kernels; 0 = dense 16-bit copies")
    ap.add_argument("--wq-stage-mb", type=int, default=0, help="staged W-only path: MB of 16-bit weights per N chunk (0 = library default)")
    ap.add_argument("--mx-bn", type=int, default=0, help="N tile of the block-scaled GEMM (0 = auto / 128 / 256)")
    ap.add_argument("--cta-group", type=int, default=0, help="GEMM CTA group (0 = auto: CTA pairs / 1 / 2)")
    ap.add_argument("--sp", action="store_true",
                    help="Ulysses sequence-parallel mode: all ranks cooperate on ONE denoising step (strong scaling); "
                         "use with --model dev --res 2048 (BASELINE.json configs[3])")
    ap.add_argument("--sp-mode", type=int, default=1, help="0 = NCCL all-to-all, 1 = peer-memory stores fused into the kernels")
    ap.add_argument("--res", type=int, default=1024, help="square resolution in pixels (--sp mode)")
    ap.add_argument("--refs", type=int, default=0, help="number of reference images (image-to-image conditioning tokens); without --sp also times the KV-cached loop")
    ap.add_argument("--lora", action="store_true", help="--sp mode: merge a rank-16 LoRA into every attention / FF linear at load time")
    ap.add_argument("--no-sp-extra", dest="sp_extra", action="store_false",
                    help="skip the extra measurements that follow the image-parallel headline (keys `sp`: Ulysses Dev 2048^2, `configs2`: Klein 9B nvfp4)")
    ap.add_argument("--sp-model", default="dev", choices=["dev", "klein9b", "klein4b"])
    ap.add_argument("--sp-res", type=int, default=2048)
    ap.add_argument("--sp-steps", type=int, default=10)
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
    from flux2b import configs   # the product's own presets / manifests; the oracle is imported by the cpu_baseline leg only

    if flux2b.device_count() < 1:
        raise SystemExit("bench.py needs an sm_100 GPU: flux2b has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    cfg = configs.PRESETS[args.model]()
    vcfg = configs.vae_small_decoder()
    if args.sp:
        run_sp(args, cfg, rank, local_rank, world, device, dist)
        return
    free0, _ = torch.cuda.mem_get_info()
    ctx = flux2b.Context(dit=cfg, vae=vcfg, device=local_rank, quant=flux2b.QUANT[args.quant],
                         options={"keep_raw_weights": 0, "native_mx": args.native_mx, "mx_bn": args.mx_bn,
                                  "gemm_cta_group": args.cta_group, "wq_inkernel": args.wq_inkernel, "wq_stage_kb": args.wq_stage_mb * 1024})
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    load_synthetic_dit(ctx, cfg, device)
    load_synthetic_vae(ctx, vcfg, device)
    ctx.finalize()
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    ctx_mem_gb = (free0 - free1) / 1e9   # resident weights of the context (DiT + VAE) after finalize

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
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "csrc_sha.txt"), "w").write(csrc_sha())   # what the profile is valid for
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
        ref_ids = torch.from_numpy(flux2b.reference_position_ids([HEIGHT // 16] * args.refs, [WIDTH // 16] * args.refs)).to(torch.int32).to(device)
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

    # Ulysses sequence parallelism in front of the driver (north_star's second split, BASELINE.json configs[3]): after the
    # image-parallel headline every run — N = 1 included, so that strong-scaling efficiency has its denominator — times one
    # Dev 32B bf16 denoising step at 2048^2 over all ranks of the job, both transports, with parity against the single-GPU step.
    sp = cfg2 = None
    if args.sp_extra and args.model == "klein4b" and args.quant == "bf16":
        ctx.close()
        del ctx
        torch.cuda.empty_cache()
        sp_args = argparse.Namespace(**vars(args))
        sp_args.quant, sp_args.native_mx = "bf16", 0
        try:
            sp = sp_measure(sp_args, configs.PRESETS[args.sp_model](), rank, local_rank, world, device, dist, args.sp_res, 0, [0, 1],
                            args.sp_steps, model_name=args.sp_model)
        except Exception as e:   # the headline must survive a failure of the extra measurement
            sp = {"error": f"{type(e).__name__}: {e}"}
        try:
            cfg2 = k9_measure(args, rank, local_rank, world, device, dist)
        except Exception as e:
            cfg2 = {"error": f"{type(e).__name__}: {e}"}

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
    traffic, traffic_src = measured_traffic("gemm") if args.quant == "bf16" and args.model == "klein4b" else (None, "not measured for this configuration")
    ncu_ms = ncu_class_ms() if args.quant == "bf16" and args.model == "klein4b" else None
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
        "mem_gb": ctx_mem_gb,
        "ms_per_step_with_kernel_events": ms_events / args.steps,
        "dit_only_steps_per_sec": n_img * NUM_STEPS / (ms_dit * 1e-3),
        "dit_tflops_per_gpu": (gemm_f + attn_f) * NUM_STEPS * args.steps / (ms_dit * 1e-3) / 1e12,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(lat_host.numel() * 4 + enc_host.numel() * 2),
                "d2h_bytes_per_step": int(HEIGHT * WIDTH * 3 + lat_host.numel() * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel (tcgen05 GEMM, all DiT linears)", "achieved": achieved,
                     "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"] if pk["tflops"] else None,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": g["bytes"] / g["launches"] if g["launches"] else None,
                     "peak_source": pk["src"], "launches": g["launches"],
                     "share_of_kernel_time": g["ms"] / total_kernel_ms if total_kernel_ms else None},
        "kernel_classes": {k: {"ms_per_image": p["ms"] / args.steps, "launches_per_image": p["launches"] / args.steps,
                               "tflops": (p["flops"] / (p["ms"] * 1e-3) / 1e12) if p["ms"] > 0 and p["flops"] else None,
                               "gbs": (p["bytes"] / (p["ms"] * 1e-3) / 1e9) if p["ms"] > 0 and p["bytes"] else None,
                               # the same class by the committed ncu launch list (kernel durations only, no event-pair overhead)
                               "ncu_ms_per_image": (ncu_ms or {}).get(k),
                               "ncu_gbs": (p["bytes"] / args.steps / ((ncu_ms or {})[k] * 1e-3) / 1e9) if (ncu_ms or {}).get(k) and p["bytes"] else None}
                           for k, p in prof.items()},
        "clocks": clocks,
    }
    if i2i is not None:
        out["i2i"] = i2i
    if sp is not None:
        out["sp"] = sp
    if cfg2 is not None:
        out["configs2"] = cfg2
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_sample(os.cpu_count() or 1)
        out["cpu_baseline"] = {"value": 1.0 / c["step_s"], "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": c["sample"], "cpu_tflops": c["tflops"]}
    emit(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
