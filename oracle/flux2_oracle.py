"""CPU oracle for the Flux2Core denoising hot path — TEST INFRASTRUCTURE ONLY.

A line-by-line restatement (PyTorch-CPU, fp32 activations like the reference) of the reference's Swift code for the
path named in SURVEY.md §8. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (libflux2b.so) never does and has no CPU path.

PARITY STATUS: **unpinned** for everything whose arithmetic lives in mlx-swift 0.31.6 (absent from /root/reference
and from this container): the reference's own tests hold no numeric golden vectors for the DiT, SDPA, RMSNorm,
conv or the quantizers. What the reference *does* pin (scheduler step counts / custom sigmas /
scaleNoise(0), position-id layouts, pack/unpack shapes, KV-extraction mask pattern, quantization table) is re-expressed
in tests/test_oracle_pins.py against this file.

Every function cites the reference file:line (relative to /root/reference/Sources/Flux2Core) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------- configs
@dataclass
class DiTConfig:
    """Flux2TransformerConfig (Configuration/Flux2Config.swift:210-329)."""
    patch_size: int = 1
    in_channels: int = 128
    out_channels: int = 128
    num_layers: int = 8
    num_single_layers: int = 48
    attention_head_dim: int = 128
    num_attention_heads: int = 48
    joint_attention_dim: int = 15360
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int, int] = (32, 32, 32, 32)
    rope_theta: float = 2000.0
    mlp_ratio: float = 3.0

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def mlp_hidden(self) -> int:
        return int(float(self.inner_dim) * self.mlp_ratio)  # Flux2TransformerBlock.swift:53


def flux2_dev() -> DiTConfig:  # Flux2Config.swift:291-300
    return DiTConfig()


def klein_4b() -> DiTConfig:  # Flux2Config.swift:302-312
    return DiTConfig(num_layers=5, num_single_layers=20, num_attention_heads=24, joint_attention_dim=7680,
                     guidance_embeds=False)


def klein_9b() -> DiTConfig:  # Flux2Config.swift:321-329
    return DiTConfig(num_layers=8, num_single_layers=24, num_attention_heads=32, joint_attention_dim=12288,
                     guidance_embeds=False)


@dataclass
class VAEConfig:
    """VAEConfig (Configuration/VAEConfig.swift:7-81)."""
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 32
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    decoder_block_out_channels: Optional[Tuple[int, ...]] = None
    layers_per_block: int = 2
    norm_eps: float = 1e-6
    norm_num_groups: int = 32

    @property
    def decoder_channels(self) -> Tuple[int, ...]:
        return self.decoder_block_out_channels or self.block_out_channels


def vae_small_decoder() -> VAEConfig:  # VAEConfig.swift:79-81
    return VAEConfig(decoder_block_out_channels=(96, 192, 384, 384))


# ----------------------------------------------------------------------------------------------- scheduler
def _f32(x) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


def compute_empirical_mu(image_seq_len: int, num_steps: int) -> float:
    """Scheduler/FlowMatchEulerScheduler.swift:9-28 (Float arithmetic)."""
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    a1, b1, a2, b2 = f(8.73809524e-05), f(1.89833333), f(0.00016927), f(0.45666666)
    n = f(float(image_seq_len))
    if image_seq_len > 4300:
        return float(a2 * n + b2)
    m_200 = a2 * n + b2
    m_10 = a1 * n + b1
    a = (m_200 - m_10) / f(190.0)
    b = m_200 - f(200.0) * a
    return float(a * f(float(num_steps)) + b)


def time_shift(mu: float, sigma: float, t: float) -> float:
    """FlowMatchEulerScheduler.swift:123-128."""
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    e = torch.exp(f(mu))
    return float(e / (e + torch.pow(f(1.0) / f(t) - f(1.0), f(sigma))))


class FlowMatchEulerScheduler:
    """Scheduler/FlowMatchEulerScheduler.swift:34-260."""

    def __init__(self, num_train_timesteps: int = 1000):
        self.num_train_timesteps = num_train_timesteps
        self.sigmas: List[float] = []
        self.timesteps: List[float] = []
        self.step_index = 0

    def set_timesteps(self, num_inference_steps: int, image_seq_len: Optional[int] = None, strength: float = 1.0) -> int:
        mu = compute_empirical_mu(image_seq_len if image_seq_len is not None else 4096, num_inference_steps)  # :67-74
        all_sigmas = []
        for i in range(num_inference_steps):  # :78-82
            s = torch.tensor(1.0, dtype=torch.float32) - torch.tensor(float(i), dtype=torch.float32) / torch.tensor(
                float(num_inference_steps), dtype=torch.float32)
            all_sigmas.append(time_shift(mu, 1.0, float(s)))  # :85-87
        all_sigmas.append(0.0)  # :90
        clamped = max(0.01, min(1.0, strength))  # :96
        init_idx = num_inference_steps - int(_f32(float(num_inference_steps)) * _f32(clamped))  # :97
        t_start = max(0, init_idx)
        self.sigmas = all_sigmas[t_start:]
        self.timesteps = [s * self.num_train_timesteps for s in self.sigmas]
        self.step_index = 0
        return t_start

    def set_custom_sigmas(self, custom: List[float]) -> None:
        """:236-260 — appends terminal 0.0 unless the last sigma is exactly 0."""
        if not custom:
            return
        s = list(custom)
        if s[-1] != 0.0:
            s.append(0.0)
        self.sigmas = s
        self.timesteps = [v * self.num_train_timesteps for v in s]
        self.step_index = 0

    @property
    def initial_sigma(self) -> float:
        return self.sigmas[0] if self.sigmas else 1.0

    def step(self, model_output: Tensor, sample: Tensor) -> Tensor:
        """:136-156."""
        if self.step_index >= len(self.sigmas) - 1:
            return sample
        dt = _f32(self.sigmas[self.step_index + 1]) - _f32(self.sigmas[self.step_index])
        self.step_index += 1
        return sample + torch.tensor(dt, dtype=torch.float32) * model_output

    @staticmethod
    def scale_noise(sample: Tensor, sigma: float, noise: Tensor) -> Tensor:
        """:195-204."""
        t = torch.tensor(sigma, dtype=torch.float32)
        return (1 - t) * sample + t * noise


# ----------------------------------------------------------------------------------------------- latent utils
def pack_patchified_to_sequence(x: Tensor) -> Tensor:
    """Pipeline/LatentUtils.swift:76-86: [B,C,H,W] -> [B,H*W,C]."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


def unpack_sequence_to_patchified(seq: Tensor, height: int, width: int) -> Tensor:
    """LatentUtils.swift:95-110."""
    B, _, C = seq.shape
    return seq.reshape(B, height // 16, width // 16, C).permute(0, 3, 1, 2)


def unpatchify_latents(x: Tensor, latent_channels: int = 32, patch: int = 2) -> Tensor:
    """LatentUtils.swift:119-142."""
    B, _, H, W = x.shape
    u = x.reshape(B, latent_channels, patch, patch, H, W).permute(0, 1, 4, 2, 5, 3)
    return u.reshape(B, latent_channels, H * patch, W * patch)


def pack_latents_to_patchified(x: Tensor, patch: int = 2) -> Tensor:
    """LatentUtils.swift:186-212."""
    B, C, H, W = x.shape
    p = x.reshape(B, C, H // patch, patch, W // patch, patch).permute(0, 2, 4, 1, 3, 5)
    p = p.reshape(B, H // patch, W // patch, C * patch * patch)
    return p.permute(0, 3, 1, 2)


def normalize_latents_bn(x: Tensor, mean: Tensor, var: Tensor, eps: float = 1e-4) -> Tensor:
    """LatentUtils.swift:460-476."""
    C = mean.shape[0]
    return (x - mean.reshape(1, C, 1, 1)) / torch.sqrt(var.reshape(1, C, 1, 1) + eps)


def denormalize_latents_bn(x: Tensor, mean: Tensor, var: Tensor, eps: float = 1e-4) -> Tensor:
    """LatentUtils.swift:483-496."""
    C = mean.shape[0]
    return x * torch.sqrt(var.reshape(1, C, 1, 1) + eps) + mean.reshape(1, C, 1, 1)


def image_position_ids(height: int, width: int, patch: int = 2) -> Tensor:
    """LatentUtils.swift:256-285: (T=0, h, w, L=0), row-major."""
    h, w = height // 8 // patch, width // 8 // patch
    hh = torch.arange(h, dtype=torch.int32).reshape(h, 1).expand(h, w).reshape(-1)
    ww = torch.arange(w, dtype=torch.int32).reshape(1, w).expand(h, w).reshape(-1)
    z = torch.zeros(h * w, dtype=torch.int32)
    return torch.stack([z, hh, ww, z], dim=1)


def text_position_ids(length: int) -> Tensor:
    """LatentUtils.swift:291-298: (0,0,0,l)."""
    z = torch.zeros(length, dtype=torch.int32)
    return torch.stack([z, z, z, torch.arange(length, dtype=torch.int32)], dim=1)


def reference_position_ids(lat_h: List[int], lat_w: List[int], scale: int = 10) -> Tensor:
    """LatentUtils.swift:324-346: reference image i gets T = scale + scale*i."""
    rows = []
    for i, (h, w) in enumerate(zip(lat_h, lat_w)):
        t = scale + scale * i
        for y in range(h):
            for x in range(w):
                rows.append([t, y, x, 0])
    return torch.tensor(rows, dtype=torch.int32).reshape(-1, 4)


def postprocess_vae_output(img: Tensor, round_nearest: bool = False) -> Tensor:
    """Pipeline/Flux2Pipeline.swift:2425-2468: (x+1)*127.5, clip, CHW->HWC, uint8 (MLX cast truncates)."""
    x = ((img[0] + 1.0) * 127.5).clamp(0, 255).permute(1, 2, 0)
    if round_nearest:
        x = torch.round(x)
    return x.to(torch.uint8)


# ----------------------------------------------------------------------------------------------- DiT pieces
def timesteps_proj(t: Tensor, num_channels: int = 256) -> Tensor:
    """Transformer/Flux2Embeddings.swift:27-44 (flipSinToCos=true, shift 0, scale 1) -> [B, 256] = [cos | sin]."""
    half = num_channels // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32)
    exponent = exponent / (float(half) - 0.0)
    emb = torch.exp(exponent)
    a = t.to(torch.float32).unsqueeze(-1) * emb.unsqueeze(0)
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return F.linear(x, w, b)


# Checker-side model of the product's OPTIONAL native block-scaled mode (option "native_mx"; not a reference behaviour: the
# reference computes x_fp32 · dequant(W)^T): the linears inside the transformer blocks see their input activation rounded
# to the 16-bit operand type and then quantised to the weight's mx / nv format. `fn` maps an fp32 tensor [..., K] to its
# fake-quantised fp32 version (oracle/quant_oracle.py: fake_quant_activation).
_BLOCK_ACT_QUANT = None


class block_activation_quant:
    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        global _BLOCK_ACT_QUANT
        self.prev, _BLOCK_ACT_QUANT = _BLOCK_ACT_QUANT, self.fn
        return self

    def __exit__(self, *exc):
        global _BLOCK_ACT_QUANT
        _BLOCK_ACT_QUANT = self.prev
        return False


def block_linear(x: Tensor, w: Tensor) -> Tensor:
    """A Linear inside a double- / single-stream block (no bias, Flux2Attention.swift:77-94)."""
    return F.linear(_BLOCK_ACT_QUANT(x) if _BLOCK_ACT_QUANT is not None else x, w)


def timestep_embedding(W: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """Flux2Embeddings.swift:74-79: Linear -> SiLU -> Linear, no bias."""
    return linear(F.silu(linear(x, W[prefix + ".linear1.weight"])), W[prefix + ".linear2.weight"])


def time_guidance_embed(W: Dict[str, Tensor], cfg: DiTConfig, timestep: Tensor, guidance: Optional[Tensor]) -> Tensor:
    """Flux2Embeddings.swift:124-141."""
    temb = timestep_embedding(W, "timeGuidanceEmbed.timestepEmbedder", timesteps_proj(timestep))
    if cfg.guidance_embeds and guidance is not None:
        temb = temb + timestep_embedding(W, "timeGuidanceEmbed.guidanceEmbedder", timesteps_proj(guidance))
    return temb


def rope_embeddings(ids: Tensor, axes_dims=(32, 32, 32, 32), theta: float = 2000.0) -> Tuple[Tensor, Tensor]:
    """Transformer/Flux2RoPE.swift:123-169 -> cos, sin [S, sum(axes)] fp32 (repeat-interleaved per axis, axes concatenated)."""
    cos_c, sin_c = [], []
    for a, dim in enumerate(axes_dims):
        pos = ids[:, a].to(torch.float32)
        freq_seq = torch.arange(0, dim, 2, dtype=torch.float32)
        inv_freq = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), freq_seq / float(dim))
        freqs = pos.unsqueeze(1) * inv_freq.unsqueeze(0)
        cos_c.append(torch.cos(freqs).repeat_interleave(2, dim=1))
        sin_c.append(torch.sin(freqs).repeat_interleave(2, dim=1))
    return torch.cat(cos_c, dim=-1), torch.cat(sin_c, dim=-1)


def rotate_half(x: Tensor) -> Tensor:
    """Flux2Attention.swift:442-461: pairs (x0,x1) -> (-x1, x0)."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    return torch.stack([-x[..., 1], x[..., 0]], dim=-1).reshape(*x.shape[:-2], -1)


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """Flux2Attention.swift:212-229: x [B,H,S,d]; cos/sin [S,d]."""
    return x * cos[None, None] + rotate_half(x) * sin[None, None]


def rms_norm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    """Flux2Attention.swift:11-26 (MLXFast.rmsNorm, fp32 accumulate)."""
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * w


def layer_norm(x: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm(eps 1e-6, affine:false): Flux2TransformerBlock.swift:56-61."""
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


def modulation(W: Dict[str, Tensor], prefix: str, temb: Tensor, num_sets: int, dim: int):
    """Flux2Modulation.swift:49-75: Linear(SiLU(temb)); per set (shift, scale, gate)."""
    allp = linear(F.silu(temb), W[prefix + ".linear.weight"])
    out = []
    for i in range(num_sets):
        s = i * dim * 3
        out.append((allp[:, s:s + dim], allp[:, s + dim:s + 2 * dim], allp[:, s + 2 * dim:s + 3 * dim]))
    return out


def apply_modulation(x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """Flux2Modulation.swift:96-112."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def apply_gate(x: Tensor, gate: Tensor) -> Tensor:
    """Flux2Modulation.swift:115-122."""
    return x * gate.unsqueeze(1)


def to_heads(x: Tensor, H: int) -> Tensor:
    """Flux2Attention.swift:199-206: [B,S,H*d] -> [B,H,S,d]."""
    B, S, _ = x.shape
    return x.reshape(B, S, H, -1).permute(0, 2, 1, 3)


def from_heads(x: Tensor) -> Tensor:
    B, H, S, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, S, H * d)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """MLXFast.scaledDotProductAttention(scale: 1/sqrt(head_dim)) — softmax in fp32."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=1.0 / math.sqrt(q.shape[-1]))


def kv_extraction_mask(text_len: int, ref_len: int, output_len: int) -> Tensor:
    """Flux2Attention.swift:422-437: additive mask, -inf where reference queries would see output keys."""
    total = text_len + ref_len + output_len
    m = torch.zeros(total, total, dtype=torch.float32)
    m[text_len:text_len + ref_len, text_len + ref_len:] = float("-inf")
    return m.reshape(1, 1, total, total)


def joint_attention(W, p, cfg: DiTConfig, img: Tensor, txt: Tensor, cos: Tensor, sin: Tensor,
                    mask: Optional[Tensor] = None, extra_kv=None, return_kv_ref: int = 0):
    """Flux2Attention.callAsFunction (Flux2Attention.swift:103-193); KV variants :245-414."""
    H = cfg.num_attention_heads
    S_txt = txt.shape[1]
    q, k, v = (to_heads(block_linear(img, W[p + n + ".weight"]), H) for n in ("attn.toQ", "attn.toK", "attn.toV"))
    aq, ak, av = (to_heads(block_linear(txt, W[p + n + ".weight"]), H) for n in ("attn.addQProj", "attn.addKProj", "attn.addVProj"))
    one = torch.ones(cfg.attention_head_dim)
    q = rms_norm(q, W.get(p + "attn.normQ.weight", one))
    k = rms_norm(k, W.get(p + "attn.normK.weight", one))
    aq = rms_norm(aq, W.get(p + "attn.normAddedQ.weight", one))
    ak = rms_norm(ak, W.get(p + "attn.normAddedK.weight", one))
    q, k = apply_rope(q, cos[S_txt:], sin[S_txt:]), apply_rope(k, cos[S_txt:], sin[S_txt:])      # :145-158
    aq, ak = apply_rope(aq, cos[:S_txt], sin[:S_txt]), apply_rope(ak, cos[:S_txt], sin[:S_txt])
    kv_ref = (k[:, :, :return_kv_ref], v[:, :, :return_kv_ref]) if return_kv_ref else None       # :292-294
    Q = torch.cat([aq, q], dim=2)                                                                 # :161-163
    if extra_kv is not None:                                                                      # :393-395 [txt | cachedRef | img]
        K = torch.cat([ak, extra_kv[0], k], dim=2)
        V = torch.cat([av, extra_kv[1], v], dim=2)
    else:
        K = torch.cat([ak, k], dim=2)
        V = torch.cat([av, v], dim=2)
    o = from_heads(sdpa(Q, K, V, mask))
    txt_o, img_o = o[:, :S_txt], o[:, S_txt:]                                                     # :178-179
    return block_linear(img_o, W[p + "attn.toOut.weight"]), block_linear(txt_o, W[p + "attn.toAddOut.weight"]), kv_ref


def feed_forward(W, p: str, x: Tensor) -> Tensor:
    """Flux2FeedForward.swift:59-67,102-108: Linear D->2Hm, split (gate, value), silu(gate)*value, Linear Hm->D."""
    h = block_linear(x, W[p + ".activation.proj.weight"])
    gate, value = h.chunk(2, dim=-1)
    return block_linear(F.silu(gate) * value, W[p + ".linearOut.weight"])


def double_block(W, i: int, cfg: DiTConfig, img, txt, img_mod, txt_mod, cos, sin, mask=None, extra_kv=None, return_kv_ref=0):
    """Flux2TransformerBlock.callAsFunction (Flux2TransformerBlock.swift:80-168); returns (txt, img)."""
    p = f"transformerBlocks.{i}."
    img_n = apply_modulation(layer_norm(img), img_mod[0][0], img_mod[0][1])
    txt_n = apply_modulation(layer_norm(txt), txt_mod[0][0], txt_mod[0][1])
    img_a, txt_a, kv = joint_attention(W, p, cfg, img_n, txt_n, cos, sin, mask, extra_kv, return_kv_ref)
    img = img + apply_gate(img_a, img_mod[0][2])
    txt = txt + apply_gate(txt_a, txt_mod[0][2])
    img_n = apply_modulation(layer_norm(img), img_mod[1][0], img_mod[1][1])
    txt_n = apply_modulation(layer_norm(txt), txt_mod[1][0], txt_mod[1][1])
    img = img + apply_gate(feed_forward(W, p + "ff", img_n), img_mod[1][2])
    txt = txt + apply_gate(feed_forward(W, p + "ffContext", txt_n), txt_mod[1][2])
    return txt, img, kv


def single_block(W, i: int, cfg: DiTConfig, x: Tensor, mod, cos, sin, mask=None, extra_kv=None, kv_slice=None, S_txt=0):
    """Flux2SingleTransformerBlock (Flux2SingleBlock.swift:59-98) + Flux2ParallelSelfAttention
    (Flux2ParallelAttention.swift:72-123; KV variants :138-269)."""
    p = f"singleTransformerBlocks.{i}."
    D, Hm, H = cfg.inner_dim, cfg.mlp_hidden, cfg.num_attention_heads
    xn = apply_modulation(layer_norm(x), mod[0][0], mod[0][1])
    proj = block_linear(xn, W[p + "attn.toQkvMlp.weight"])                       # q | k | v | gate | up  (:83-87)
    q, k, v = (to_heads(proj[..., j * D:(j + 1) * D], H) for j in range(3))
    gate, up = proj[..., 3 * D:3 * D + Hm], proj[..., 3 * D + Hm:]
    one = torch.ones(cfg.attention_head_dim)
    q = rms_norm(q, W.get(p + "attn.normQ.weight", one))
    k = rms_norm(k, W.get(p + "attn.normK.weight", one))
    if extra_kv is not None:
        # cached pass: queries [txt | img] use the rope rows of [txt | img]; keys [txt | cachedRef | img] (:247-253)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        K = torch.cat([k[:, :, :S_txt], extra_kv[0], k[:, :, S_txt:]], dim=2)
        V = torch.cat([v[:, :, :S_txt], extra_kv[1], v[:, :, S_txt:]], dim=2)
        kv = None
    else:
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        K, V = k, v
        kv = (k[:, :, kv_slice[0]:kv_slice[1]], v[:, :, kv_slice[0]:kv_slice[1]]) if kv_slice else None
    attn = from_heads(sdpa(q, K, V, mask))
    out = block_linear(torch.cat([attn, F.silu(gate) * up], dim=-1), W[p + "attn.toOut.weight"])   # :116-122
    return x + apply_gate(out, mod[0][2]), kv                                             # Flux2SingleBlock.swift:91-97


def dit_forward(W: Dict[str, Tensor], cfg: DiTConfig, hidden: Tensor, enc: Tensor, timestep: Tensor,
                guidance: Optional[Tensor], img_ids: Tensor, txt_ids: Tensor, record: Optional[list] = None,
                kv_mode: int = 0, ref_hidden: Optional[Tensor] = None, ref_ids: Optional[Tensor] = None,
                kv_cache: Optional[list] = None):
    """Flux2Transformer2DModel.callAsFunction (Transformer/Flux2Transformer.swift:123-327);
    kv_mode 1 = forwardKVExtract (:346-457, tokens [txt | ref | img]), 2 = forwardKVCached (:459-546)."""
    D = cfg.inner_dim
    S_txt = enc.shape[1]
    if kv_mode == 1:
        S_ref = ref_hidden.shape[1]
        img_in = torch.cat([ref_hidden, hidden], dim=1)          # :364
        ids = torch.cat([txt_ids, ref_ids, img_ids], dim=0)      # :372
    else:
        S_ref = 0
        img_in = hidden
        ids = torch.cat([txt_ids, img_ids], dim=0)               # :153
    img = linear(img_in.to(torch.float32), W["xEmbedder.weight"])              # :137
    txt = linear(enc.to(torch.float32), W["contextEmbedder.weight"])           # :138
    temb = time_guidance_embed(W, cfg, timestep * 1000.0, guidance * 1000.0 if guidance is not None else None)  # :145-149
    cos, sin = rope_embeddings(ids, cfg.axes_dims_rope, cfg.rope_theta)
    img_mod = modulation(W, "doubleStreamModulationImg", temb, 2, D)           # :160-161
    txt_mod = modulation(W, "doubleStreamModulationTxt", temb, 2, D)
    mask = kv_extraction_mask(S_txt, S_ref, hidden.shape[1]) if kv_mode == 1 else None
    new_cache = []
    layer = 0
    for i in range(cfg.num_layers):                                            # :168
        extra = kv_cache[layer] if kv_mode == 2 else None
        txt, img, kv = double_block(W, i, cfg, img, txt, img_mod, txt_mod, cos, sin, mask, extra, S_ref if kv_mode == 1 else 0)
        new_cache.append(kv)
        layer += 1
        if record is not None:
            record.append(torch.cat([txt, img], dim=1)[0].clone())
    x = torch.cat([txt, img], dim=1)                                           # :252
    s_mod = modulation(W, "singleStreamModulation", temb, 1, D)                # :256
    for i in range(cfg.num_single_layers):                                     # :259
        extra = kv_cache[layer] if kv_mode == 2 else None
        x, kv = single_block(W, i, cfg, x, s_mod, cos, sin, mask, extra,
                             (S_txt, S_txt + S_ref) if kv_mode == 1 else None, S_txt)
        new_cache.append(kv)
        layer += 1
        if record is not None:
            record.append(x[0].clone())
    img = x[:, S_txt + S_ref:, :]                                              # :316
    params = linear(F.silu(temb), W["normOut.linear.weight"])                  # Flux2Modulation.swift:142-155
    scale, shift = params[:, :D], params[:, D:]                                # scale first (:146-148)
    img = apply_modulation(layer_norm(img), shift, scale)
    out = linear(img, W["projOut.weight"])                                     # :324
    if kv_mode == 1:
        return out, new_cache
    return out


# ----------------------------------------------------------------------------------------------- random-init weights
def _uniform_linear(gen: torch.Generator, out_f: int, in_f: int, round_to: Optional[torch.dtype]) -> Tensor:
    """MLX Linear default init U(-1/sqrt(in), 1/sqrt(in)), optionally rounded to a 16-bit grid so that the
    f16 weights the reference would hold and the 16-bit weights of the device path are the same numbers."""
    k = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float32) * 2 - 1) * k
    return w.to(round_to).to(torch.float32) if round_to is not None else w


def dit_weight_shapes(cfg: DiTConfig) -> Dict[str, Tuple[int, int]]:
    D, Hm = cfg.inner_dim, cfg.mlp_hidden
    s = {
        "xEmbedder.weight": (D, cfg.in_channels),
        "contextEmbedder.weight": (D, cfg.joint_attention_dim),
        "timeGuidanceEmbed.timestepEmbedder.linear1.weight": (D, 256),
        "timeGuidanceEmbed.timestepEmbedder.linear2.weight": (D, D),
        "doubleStreamModulationImg.linear.weight": (6 * D, D),
        "doubleStreamModulationTxt.linear.weight": (6 * D, D),
        "singleStreamModulation.linear.weight": (3 * D, D),
        "normOut.linear.weight": (2 * D, D),
        "projOut.weight": (cfg.out_channels, D),
    }
    if cfg.guidance_embeds:
        s["timeGuidanceEmbed.guidanceEmbedder.linear1.weight"] = (D, 256)
        s["timeGuidanceEmbed.guidanceEmbedder.linear2.weight"] = (D, D)
    for i in range(cfg.num_layers):
        p = f"transformerBlocks.{i}."
        for n in ("attn.toQ", "attn.toK", "attn.toV", "attn.addQProj", "attn.addKProj", "attn.addVProj", "attn.toOut", "attn.toAddOut"):
            s[p + n + ".weight"] = (D, D)
        for ff in ("ff", "ffContext"):
            s[p + ff + ".activation.proj.weight"] = (2 * Hm, D)
            s[p + ff + ".linearOut.weight"] = (D, Hm)
    for i in range(cfg.num_single_layers):
        p = f"singleTransformerBlocks.{i}."
        s[p + "attn.toQkvMlp.weight"] = (3 * D + 2 * Hm, D)
        s[p + "attn.toOut.weight"] = (D, D + Hm)
    return s


def random_dit_weights(cfg: DiTConfig, seed: int = 0, round_to: Optional[torch.dtype] = torch.bfloat16,
                       norm_weights: bool = True) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    W = {k: _uniform_linear(gen, o, i, round_to) for k, (o, i) in dit_weight_shapes(cfg).items()}
    if norm_weights:  # RMSNorm weights: ones in the reference init; perturbed here so that a wrong weight shows up
        def nw():
            v = 1.0 + 0.1 * torch.randn(cfg.attention_head_dim, generator=gen)
            return v.to(torch.float32)
        for i in range(cfg.num_layers):
            for n in ("normQ", "normK", "normAddedQ", "normAddedK"):
                W[f"transformerBlocks.{i}.attn.{n}.weight"] = nw()
        for i in range(cfg.num_single_layers):
            for n in ("normQ", "normK"):
                W[f"singleTransformerBlocks.{i}.attn.{n}.weight"] = nw()
    return W


# ----------------------------------------------------------------------------------------------- VAE decoder
def group_norm_nhwc(x: Tensor, w: Tensor, b: Tensor, groups: int, eps: float) -> Tensor:
    """VAE/ResnetBlock.swift:24-54: fp32 statistics over (H, W, C/G), affine after."""
    B, H, W, C = x.shape
    r = x.to(torch.float32).reshape(B, H, W, groups, C // groups)
    mean = r.mean(dim=(1, 2, 4), keepdim=True)
    var = ((r - mean) ** 2).mean(dim=(1, 2, 4), keepdim=True)
    n = ((r - mean) / torch.sqrt(var + eps)).reshape(B, H, W, C)
    return n * w.reshape(1, 1, 1, C) + b.reshape(1, 1, 1, C)


def conv2d_nhwc(x: Tensor, w_ohwi: Tensor, b: Optional[Tensor], padding: int) -> Tensor:
    """MLX Conv2d on NHWC input with OHWI weights (WeightLoader.swift:496-498), zero padding."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w_ohwi.permute(0, 3, 1, 2), b, padding=padding)
    return y.permute(0, 2, 3, 1)


def resnet_block(W, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    """ResnetBlock.swift:168-186."""
    h = F.silu(group_norm_nhwc(x, W[p + ".norm1.weight"], W[p + ".norm1.bias"], groups, eps))
    h = conv2d_nhwc(h, W[p + ".conv1.weight"], W[p + ".conv1.bias"], 1)
    h = F.silu(group_norm_nhwc(h, W[p + ".norm2.weight"], W[p + ".norm2.bias"], groups, eps))
    h = conv2d_nhwc(h, W[p + ".conv2.weight"], W[p + ".conv2.bias"], 1)
    sc = conv2d_nhwc(x, W[p + ".convShortcut.weight"], W[p + ".convShortcut.bias"], 0) if (p + ".convShortcut.weight") in W else x
    return h + sc


def vae_attention_block(W, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    """ResnetBlock.swift:281-313: single head, explicit softmax(QK^T/sqrt(C)) V."""
    B, H, Wd, C = x.shape
    h = group_norm_nhwc(x, W[p + ".groupNorm.weight"], W[p + ".groupNorm.bias"], groups, eps).reshape(B, H * Wd, C)
    q = linear(h, W[p + ".toQ.weight"], W[p + ".toQ.bias"])
    k = linear(h, W[p + ".toK.weight"], W[p + ".toK.bias"]).transpose(1, 2)
    v = linear(h, W[p + ".toV.weight"], W[p + ".toV.bias"])
    a = torch.softmax(torch.matmul(q, k) * (1.0 / math.sqrt(float(C))), dim=-1)
    o = linear(torch.matmul(a, v), W[p + ".toOut.weight"], W[p + ".toOut.bias"])
    return o.reshape(B, H, Wd, C) + x


def upsample2d(W, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.swift:229-253: nearest x2 then conv3x3 pad 1."""
    x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    return conv2d_nhwc(x, W[p + ".conv.weight"], W[p + ".conv.bias"], 1)


def vae_decode(W: Dict[str, Tensor], cfg: VAEConfig, z: Tensor) -> Tensor:
    """AutoencoderKLFlux2.decode (VAE/AutoencoderKL.swift:129-143) + VAEDecoder (VAE/VAEDecoder.swift:91-121).
    z [B, 32, h, w] NCHW -> [B, 3, 8h, 8w] NCHW."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    x = z.to(torch.float32).permute(0, 2, 3, 1)
    x = conv2d_nhwc(x, W["postQuantConv.weight"], W["postQuantConv.bias"], 0)
    x = conv2d_nhwc(x, W["decoder.convIn.weight"], W["decoder.convIn.bias"], 1)
    x = resnet_block(W, "decoder.midBlock.0", x, g, eps)
    x = vae_attention_block(W, "decoder.midBlock.1", x, g, eps)
    x = resnet_block(W, "decoder.midBlock.2", x, g, eps)
    for i in range(4):
        for j in range(cfg.layers_per_block + 1):
            x = resnet_block(W, f"decoder.upBlocks.{i}.0.{j}", x, g, eps)
        if i < 3:
            x = upsample2d(W, f"decoder.upBlocks.{i}.1", x)
    x = F.silu(group_norm_nhwc(x, W["decoder.convNormOut.weight"], W["decoder.convNormOut.bias"], g, eps))
    x = conv2d_nhwc(x, W["decoder.convOut.weight"], W["decoder.convOut.bias"], 1)
    return x.permute(0, 3, 1, 2)


def downsample2d(W, p: str, x: Tensor) -> Tensor:
    """Downsample2D (ResnetBlock.swift:189-213): zero-pad bottom / right by one pixel only, conv3x3 stride 2 without padding."""
    x = F.pad(x, (0, 0, 0, 1, 0, 1))   # NHWC: (C: 0,0) (W: 0,1) (H: 0,1)
    y = F.conv2d(x.permute(0, 3, 1, 2), W[p + ".conv.weight"].permute(0, 3, 1, 2), W[p + ".conv.bias"], stride=2, padding=0)
    return y.permute(0, 2, 3, 1)


def vae_encode_moments(W: Dict[str, Tensor], cfg: VAEConfig, img: Tensor) -> Tensor:
    """VAEEncoder.callAsFunction (VAE/VAEEncoder.swift:85-115) + quantConv (VAE/AutoencoderKL.swift:94-99).
    img [B, 3, H, W] NCHW in [-1, 1] -> posterior moments [B, 2 * latent, H/8, W/8] NCHW (mean | logvar)."""
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    x = img.to(torch.float32).permute(0, 2, 3, 1)
    x = conv2d_nhwc(x, W["encoder.convIn.weight"], W["encoder.convIn.bias"], 1)
    for i in range(len(cfg.block_out_channels)):
        for j in range(cfg.layers_per_block):
            x = resnet_block(W, f"encoder.downBlocks.{i}.0.{j}", x, g, eps)
        if i < len(cfg.block_out_channels) - 1:
            x = downsample2d(W, f"encoder.downBlocks.{i}.1", x)
    x = resnet_block(W, "encoder.midBlock.0", x, g, eps)
    x = vae_attention_block(W, "encoder.midBlock.1", x, g, eps)
    x = resnet_block(W, "encoder.midBlock.2", x, g, eps)
    x = F.silu(group_norm_nhwc(x, W["encoder.convNormOut.weight"], W["encoder.convNormOut.bias"], g, eps))
    x = conv2d_nhwc(x, W["encoder.convOut.weight"], W["encoder.convOut.bias"], 1)
    if "quantConv.weight" in W:
        x = conv2d_nhwc(x, W["quantConv.weight"], W["quantConv.bias"], 0)
    return x.permute(0, 3, 1, 2)


def vae_encode(W: Dict[str, Tensor], cfg: VAEConfig, img: Tensor, noise: Optional[Tensor] = None) -> Tensor:
    """AutoencoderKLFlux2.encode (VAE/AutoencoderKL.swift:90-127): mean (samplePosterior false) or mean + exp(logvar / 2) * noise;
    no scaling factor, no BatchNorm (:113-123)."""
    h = vae_encode_moments(W, cfg, img)
    L = cfg.latent_channels
    mean, logvar = h[:, :L], h[:, L:]
    return mean if noise is None else mean + torch.exp(0.5 * logvar) * noise


def encode_image_to_packed_sequence(W: Dict[str, Tensor], cfg: VAEConfig, img: Tensor) -> Tensor:
    """encodeImageToPackedSequence (Flux2Pipeline+ChainHelpers.swift:75-101) = per-image body of encodeReferenceImages
    (Flux2Pipeline.swift:2196-2213), after image preprocessing."""
    pat = pack_latents_to_patchified(vae_encode(W, cfg, img))
    pat = normalize_latents_bn(pat, W["latentBatchNorm.runningMean"], W["latentBatchNorm.runningVar"], 1e-4)
    return pack_patchified_to_sequence(pat)


def random_vae_weights(cfg: VAEConfig, seed: int = 1, round_to: Optional[torch.dtype] = torch.float16,
                       encoder: bool = False) -> Dict[str, Tensor]:
    """Synthetic decoder (+ optionally encoder) weights: fan-in-scaled uniform convs / linears, GN gamma ~ 1, beta ~ 0,
    BN mean 0 var 1 (+noise). The encoder tensors are drawn after all decoder tensors, so decoder fixtures do not change."""
    gen = torch.Generator().manual_seed(seed)
    W: Dict[str, Tensor] = {}

    def rnd(t):
        return t.to(round_to).to(torch.float32) if round_to is not None else t

    def conv(name, co, ci, k):
        bound = 1.0 / math.sqrt(ci * k * k)
        W[name + ".weight"] = rnd((torch.rand(co, k, k, ci, generator=gen) * 2 - 1) * bound)
        W[name + ".bias"] = (torch.rand(co, generator=gen) * 2 - 1) * bound

    def lin(name, co, ci):
        bound = 1.0 / math.sqrt(ci)
        W[name + ".weight"] = rnd((torch.rand(co, ci, generator=gen) * 2 - 1) * bound)
        W[name + ".bias"] = (torch.rand(co, generator=gen) * 2 - 1) * bound

    def norm(name, c):
        W[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=gen)
        W[name + ".bias"] = 0.1 * torch.randn(c, generator=gen)

    def resnet(name, ci, co):
        norm(name + ".norm1", ci); conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co); conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".convShortcut", co, ci, 1)

    ch = cfg.decoder_channels
    L = cfg.latent_channels
    conv("postQuantConv", L, L, 1)
    conv("decoder.convIn", ch[3], L, 3)
    resnet("decoder.midBlock.0", ch[3], ch[3])
    norm("decoder.midBlock.1.groupNorm", ch[3])
    for n in ("toQ", "toK", "toV", "toOut"):
        lin("decoder.midBlock.1." + n, ch[3], ch[3])
    resnet("decoder.midBlock.2", ch[3], ch[3])
    prev = ch[3]
    for i in range(4):
        co = ch[3 - i]
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.upBlocks.{i}.0.{j}", prev if j == 0 else co, co)
        prev = co
        if i < 3:
            conv(f"decoder.upBlocks.{i}.1.conv", co, co, 3)
    norm("decoder.convNormOut", ch[0])
    conv("decoder.convOut", cfg.out_channels, ch[0], 3)
    W["latentBatchNorm.runningMean"] = 0.1 * torch.randn(128, generator=gen)
    W["latentBatchNorm.runningVar"] = 1.0 + 0.1 * torch.rand(128, generator=gen)
    if encoder:
        ec = cfg.block_out_channels
        conv("encoder.convIn", ec[0], cfg.in_channels, 3)
        prev = ec[0]
        for i, co in enumerate(ec):
            for j in range(cfg.layers_per_block):
                resnet(f"encoder.downBlocks.{i}.0.{j}", prev, co)
                prev = co
            if i < len(ec) - 1:
                conv(f"encoder.downBlocks.{i}.1.conv", co, co, 3)
        resnet("encoder.midBlock.0", ec[-1], ec[-1])
        norm("encoder.midBlock.1.groupNorm", ec[-1])
        for n in ("toQ", "toK", "toV", "toOut"):
            lin("encoder.midBlock.1." + n, ec[-1], ec[-1])
        resnet("encoder.midBlock.2", ec[-1], ec[-1])
        norm("encoder.convNormOut", ec[-1])
        conv("encoder.convOut", 2 * L, ec[-1], 3)
        conv("quantConv", 2 * L, 2 * L, 1)
    return W


# ----------------------------------------------------------------------------------------------- pipeline tail / loop
def latents_to_image(W_vae, vcfg: VAEConfig, seq: Tensor, height: int, width: int) -> Tensor:
    """Flux2Pipeline.swift:2059-2098: unpack -> BN denorm -> unpatchify -> decode."""
    pat = unpack_sequence_to_patchified(seq, height, width)
    pat = denormalize_latents_bn(pat, W_vae["latentBatchNorm.runningMean"], W_vae["latentBatchNorm.runningVar"], 1e-4)
    return vae_decode(W_vae, vcfg, unpatchify_latents(pat))


def denoise(W, cfg: DiTConfig, latents: Tensor, enc: Tensor, sigmas: List[float], height: int, width: int,
            guidance: Optional[float] = None, enc_uncond: Optional[Tensor] = None, cfg_scale: float = 1.0,
            hook=None, ref_latents: Optional[Tensor] = None, ref_ids: Optional[Tensor] = None, kv_cache: bool = False) -> Tensor:
    """The per-step body of Flux2Pipeline.generateWithResult (Flux2Pipeline.swift:1933-2001; I2I :1696-1767;
    kv_cache=True: the klein-9b-kv loop :1565-1644 — forwardKVExtract at step 0, forwardKVCached afterwards)."""
    txt_ids = text_position_ids(enc.shape[1])
    img_ids = image_position_ids(height, width)
    S_img = latents.shape[1]
    x = latents.clone()
    g = torch.tensor([guidance], dtype=torch.float32) if guidance is not None else None
    for i in range(len(sigmas) - 1):
        t = torch.tensor([sigmas[i]], dtype=torch.float32)
        if kv_cache and ref_latents is not None:
            if i == 0:
                pred, cache = dit_forward(W, cfg, x, enc, t, g, img_ids, txt_ids, kv_mode=1, ref_hidden=ref_latents, ref_ids=ref_ids)
            else:
                pred = dit_forward(W, cfg, x, enc, t, g, img_ids, txt_ids, kv_mode=2, kv_cache=cache)
            dt = _f32(sigmas[i + 1]) - _f32(sigmas[i])
            x = x + torch.tensor(dt, dtype=torch.float32) * pred
            if hook is not None:
                x = hook(i, len(sigmas) - 1, sigmas[i], sigmas[i + 1], x)
            continue
        if ref_latents is not None:
            hid = torch.cat([x, ref_latents], dim=1)                 # [output | refs] (:1703)
            ids = torch.cat([img_ids, ref_ids], dim=0)               # (:1504)
        else:
            hid, ids = x, img_ids
        pred = dit_forward(W, cfg, hid, enc, t, g, ids, txt_ids)[:, :S_img]   # (:1743)
        if enc_uncond is not None:
            # the negative prompt's ids follow its own length (uncondTextIds, Flux2Pipeline.swift:1687-1694, 1919-1925)
            un_ids = text_position_ids(enc_uncond.shape[1])
            un = dit_forward(W, cfg, hid, enc_uncond, t, g, ids, un_ids)[:, :S_img]
            pred = un + cfg_scale * (pred - un)                      # (:1970)
        dt = _f32(sigmas[i + 1]) - _f32(sigmas[i])
        x = x + torch.tensor(dt, dtype=torch.float32) * pred         # FlowMatchEulerScheduler.swift:150-151
        if hook is not None:
            x = hook(i, len(sigmas) - 1, sigmas[i], sigmas[i + 1], x)
    return x


def repaint_blend(x: Tensor, x0: Tensor, eps: Tensor, mask: Tensor, sigma_next: float) -> Tensor:
    """Flux2Chains/Flux2MaskedInpaintingChain.swift:399-403."""
    known = (1 - sigma_next) * x0 + sigma_next * eps
    return (1 - mask) * known + mask * x


def lora_merge(weight: Tensor, A: Tensor, B: Tensor, scale: float, dtype: torch.dtype) -> Tensor:
    """Loading/WeightLoader.swift:825-838: everything in the weight dtype (matmul accumulates in fp32 inside MLX)."""
    Ad, Bd = A.to(dtype).to(torch.float32), B.to(dtype).to(torch.float32)
    r = A.shape[0]
    acc = torch.zeros(B.shape[0], A.shape[1], dtype=torch.float32)
    for j in range(r):  # sequential fp32 accumulation over the rank (restated; MLX's blocked order is unpinned)
        acc = acc + Bd[:, j:j + 1] * Ad[j:j + 1, :]
    ba = acc.to(dtype).to(torch.float32)
    s = torch.tensor(scale, dtype=torch.float32).to(dtype).to(torch.float32)
    delta = (s * ba).to(dtype).to(torch.float32)
    return (weight.to(dtype).to(torch.float32) + delta).to(dtype)


# ----------------------------------------------------------------------------------------------- text encoder
# Cites below are relative to /root/reference/Sources/FluxTextEncoders. PARITY STATUS: the reference holds no numeric vectors for its
# text encoders and RMSNorm / RoPE / SDPA arithmetic lives in mlx-swift (unpinned); the STRUCTURE restated here is pinned against
# Hugging Face transformers' Qwen3Model / MistralModel — the modules the Swift files are ports of — in tests/test_te_hf_pin_cpu.py. What the reference does fix — layer
# indexing, padding side, mask values, pair layout of the rotation, GQA head mapping — is restated here and self-tested in
# tests/test_oracle_pins.py.
@dataclass
class TEConfig:
    """Qwen3TextConfig (Configuration/Qwen3Configuration.swift:16-130) / MistralTextConfig."""
    vocab_size: int = 151_936
    hidden_size: int = 2560
    intermediate_size: int = 9216
    num_layers: int = 36
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    qk_norm: bool = True           # Qwen3: q_norm / k_norm (Qwen3Attention.swift:58-61); Mistral: none
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    max_position_embeddings: int = 0


def qwen3_4b() -> TEConfig:  # Qwen3Configuration.swift:74-90 (head_dim 128 is what the checkpoint's config.json carries)
    return TEConfig()


def qwen3_8b() -> TEConfig:  # Qwen3Configuration.swift:93-109
    return TEConfig(hidden_size=4096, intermediate_size=12288)


KLEIN_HIDDEN_STATE_LAYERS = (9, 18, 27)   # Embeddings/KleinConfig.swift:28-31
FLUX_HIDDEN_STATE_LAYERS = (10, 20, 30)   # Embeddings/EmbeddingExtractor.swift (FluxConfig.hiddenStateLayers)


def te_weight_shapes(cfg: TEConfig, layers: Optional[int] = None) -> Dict[str, Tuple[int, ...]]:
    """Module paths of Qwen3ForCausalLM (Model/Qwen3/Qwen3Model.swift:33-55, Qwen3DecoderLayer.swift:14-18,
    Qwen3Attention.swift:54-62, Qwen3MLP.swift:19-21) = the HF checkpoint keys."""
    Hd, I = cfg.hidden_size, cfg.intermediate_size
    s: Dict[str, Tuple[int, ...]] = {"model.embed_tokens.weight": (cfg.vocab_size, Hd), "model.norm.weight": (Hd,)}
    for i in range(cfg.num_layers if layers is None else layers):
        p = f"model.layers.{i}."
        s[p + "self_attn.q_proj.weight"] = (cfg.num_heads * cfg.head_dim, Hd)
        s[p + "self_attn.k_proj.weight"] = (cfg.num_kv_heads * cfg.head_dim, Hd)
        s[p + "self_attn.v_proj.weight"] = (cfg.num_kv_heads * cfg.head_dim, Hd)
        s[p + "self_attn.o_proj.weight"] = (Hd, cfg.num_heads * cfg.head_dim)
        s[p + "mlp.gate_proj.weight"] = (I, Hd)
        s[p + "mlp.up_proj.weight"] = (I, Hd)
        s[p + "mlp.down_proj.weight"] = (Hd, I)
        s[p + "input_layernorm.weight"] = (Hd,)
        s[p + "post_attention_layernorm.weight"] = (Hd,)
        if cfg.qk_norm:
            s[p + "self_attn.q_norm.weight"] = (cfg.head_dim,)
            s[p + "self_attn.k_norm.weight"] = (cfg.head_dim,)
    return s


def random_te_weights(cfg: TEConfig, seed: int = 2, layers: Optional[int] = None,
                      round_to: Optional[torch.dtype] = torch.bfloat16) -> Dict[str, Tensor]:
    """Random-init text encoder: Linear U(-1/sqrt(in), 1/sqrt(in)) (MLX default), Embedding N(0, 1) scaled to the residual
    magnitude of a trained model's embedding (std 0.05), norm weights perturbed around one so that a wrong weight shows."""
    gen = torch.Generator().manual_seed(seed)
    W: Dict[str, Tensor] = {}
    for k, shp in te_weight_shapes(cfg, layers).items():
        if len(shp) == 1:
            W[k] = (1.0 + 0.1 * torch.randn(shp[0], generator=gen)).to(torch.float32)
        elif k == "model.embed_tokens.weight":
            w = torch.randn(shp[0], shp[1], generator=gen) * 0.05
            W[k] = w.to(round_to).to(torch.float32) if round_to is not None else w
        else:
            W[k] = _uniform_linear(gen, shp[0], shp[1], round_to)
    return W


def te_rope_half(x: Tensor, base: float, offset: int = 0) -> Tensor:
    """MLXFast.RoPE(traditional: false, scale 1) as used by Qwen3RoPE / MistralRoPE (Model/Qwen3/Qwen3Attention.swift:29-35,
    Model/MistralAttention.swift:336-357): x [B, H, S, hd]; element j < hd/2 pairs with j + hd/2, angle = pos * base^(-j/(hd/2))
    (upstream kernel: inv_freq = exp2(-(j / (hd/2)) * log2(base)), fp32)."""
    hd = x.shape[-1]
    half = hd // 2
    j = torch.arange(half, dtype=torch.float32)
    inv_freq = torch.exp2(-(j / half) * math.log2(base)).to(torch.float32)
    pos = torch.arange(offset, offset + x.shape[-2], dtype=torch.float32)
    ang = pos[:, None] * inv_freq[None, :]
    cos, sin = torch.cos(ang), torch.sin(ang)
    x1, x2 = x[..., :half], x[..., half:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)


def te_causal_mask(seq_len: int, attention_mask: Optional[Tensor]) -> Tensor:
    """Qwen3Model.createCausalMask (Model/Qwen3/Qwen3Model.swift:196-231; MistralModel.swift:150-194): 0 where j <= i else
    -inf, plus -1e9 on padded keys (attention_mask == 0), fp32, shape [1, 1, S, S]."""
    i = torch.arange(seq_len, dtype=torch.float32)[:, None]
    j = torch.arange(seq_len, dtype=torch.float32)[None, :]
    mask = torch.where(j <= i, torch.tensor(0.0), torch.tensor(-float("inf")))
    mask = mask.reshape(1, 1, seq_len, seq_len)
    if attention_mask is not None:
        pad = torch.where(attention_mask.reshape(1, -1) == 1, torch.tensor(0.0), torch.tensor(-1e9)).reshape(1, 1, 1, seq_len)
        mask = mask + pad
    return mask


def _rnd(x: Tensor, dt: Optional[torch.dtype]) -> Tensor:
    return x if dt is None else x.to(dt).to(torch.float32)


def te_attention(W, p: str, cfg: TEConfig, x: Tensor, mask: Tensor, operand_dtype: Optional[torch.dtype] = None) -> Tensor:
    """Qwen3Attention.callAsFunction (Model/Qwen3/Qwen3Attention.swift:92-163); with qk_norm False it is
    MistralAttention.callAsFunction (Model/MistralAttention.swift:393-474; its Llama-4 query scale is 1 for positions below
    original_max_position_embeddings).

    operand_dtype (checker-side model, not a reference behaviour): round to that 16-bit type exactly where the device path
    stores tensor-core operands (q / k after RoPE, v, the unnormalised softmax numerator, the attention output). The
    reference itself runs the whole encoder in the checkpoint's 16-bit dtype, i.e. it rounds at MORE points than this."""
    B, S, _ = x.shape
    Hq, Hkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    q = linear(x, W[p + "q_proj.weight"]).reshape(B, S, Hq, hd)
    k = linear(x, W[p + "k_proj.weight"]).reshape(B, S, Hkv, hd)
    v = _rnd(linear(x, W[p + "v_proj.weight"]), operand_dtype).reshape(B, S, Hkv, hd)
    if cfg.qk_norm:  # per head, BEFORE RoPE (:108-111)
        q = rms_norm(q, W[p + "q_norm.weight"], cfg.rms_norm_eps)
        k = rms_norm(k, W[p + "k_norm.weight"], cfg.rms_norm_eps)
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    q = _rnd(te_rope_half(q, cfg.rope_theta), operand_dtype)
    k = _rnd(te_rope_half(k, cfg.rope_theta), operand_dtype)
    rep = Hq // Hkv  # query head h uses kv head h // rep (:133-145: expand axis 2, broadcast, reshape)
    k = k[:, :, None].expand(B, Hkv, rep, S, hd).reshape(B, Hq, S, hd)
    v = v[:, :, None].expand(B, Hkv, rep, S, hd).reshape(B, Hq, S, hd)
    scores = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)) + mask   # additive mask on the scaled scores, fp32
    if operand_dtype is None:
        o = torch.softmax(scores, dim=-1) @ v
    else:
        e = torch.exp(scores - scores.max(dim=-1, keepdim=True).values)
        o = _rnd((_rnd(e, operand_dtype) @ v) / e.sum(dim=-1, keepdim=True), operand_dtype)
    o = o.transpose(1, 2).reshape(B, S, Hq * hd)
    return linear(o, W[p + "o_proj.weight"])


def te_mlp(W, p: str, x: Tensor, operand_dtype: Optional[torch.dtype] = None) -> Tensor:
    """Qwen3MLP (Model/Qwen3/Qwen3MLP.swift:42-47): down(silu(gate(x)) * up(x))."""
    a = F.silu(linear(x, W[p + "gate_proj.weight"])) * linear(x, W[p + "up_proj.weight"])
    return linear(_rnd(a, operand_dtype), W[p + "down_proj.weight"])


def te_decoder_layer(W, i: int, cfg: TEConfig, h: Tensor, mask: Tensor, operand_dtype: Optional[torch.dtype] = None) -> Tensor:
    """Qwen3DecoderLayer (Model/Qwen3/Qwen3DecoderLayer.swift:32-48)."""
    p = f"model.layers.{i}."
    x = _rnd(rms_norm(h, W[p + "input_layernorm.weight"], cfg.rms_norm_eps), operand_dtype)
    h = h + te_attention(W, p + "self_attn.", cfg, x, mask, operand_dtype)
    x = _rnd(rms_norm(h, W[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps), operand_dtype)
    return h + te_mlp(W, p + "mlp.", x, operand_dtype)


def te_hidden_states(W, cfg: TEConfig, input_ids: Tensor, attention_mask: Optional[Tensor], layer_indices,
                     operand_dtype: Optional[torch.dtype] = None) -> Tensor:
    """Qwen3Model.forwardWithHiddenStates (Model/Qwen3/Qwen3Model.swift:104-191) + the concatenation of
    KleinEmbeddingExtractor.extractKleinEmbeddings (Embeddings/KleinEmbeddingExtractor.swift:98-121): index 0 = embedding
    output, i = output of decoder layer i (1-based), num_layers = after the final norm. input_ids [1, S] -> [1, S, n * hidden]."""
    h = W["model.embed_tokens.weight"][input_ids.long()]
    mask = te_causal_mask(input_ids.shape[1], attention_mask)
    want = set(int(x) for x in layer_indices)
    got: Dict[int, Tensor] = {}
    if 0 in want:
        got[0] = h
    for i in range(max(want)):
        h = te_decoder_layer(W, i, cfg, h, mask, operand_dtype)
        if (i + 1) in want and (i + 1) < cfg.num_layers:
            got[i + 1] = h
    if cfg.num_layers in want:
        got[cfg.num_layers] = rms_norm(h, W["model.norm.weight"], cfg.rms_norm_eps)
    return torch.cat([got[int(i)] for i in layer_indices], dim=-1)


def llama4_attention_scale(start: int, stop: int, beta: float, max_position_embeddings: int) -> Tensor:
    """getLlama4AttentionScale (Model/MistralAttention.swift:15-32): 1 + beta * log(1 + floor(pos / max_pos)), applied to the rotated
    queries (:422-432). It is exactly 1 for every position below original_max_position_embeddings, which is why the device path
    refuses longer inputs instead of computing it (512-token prompts never get there)."""
    pos = torch.arange(start, stop, dtype=torch.float32)
    return 1.0 + beta * torch.log(1.0 + torch.floor(pos / float(max_position_embeddings)))


def te_pad_tokens(token_ids: List[int], max_length: int, pad_id: int, side: str) -> Tuple[Tensor, Tensor]:
    """Truncate + pad + mask as the extractors do: Klein RIGHT-pads with <|endoftext|> 151643
    (Embeddings/KleinEmbeddingExtractor.swift:69-95), Dev LEFT-pads (Embeddings/EmbeddingExtractor.swift:221-248)."""
    ids = list(token_ids[:max_length])
    n = len(ids)
    pad = [pad_id] * (max_length - n)
    if side == "right":
        ids, m = ids + pad, [1] * n + [0] * len(pad)
    else:
        ids, m = pad + ids, [0] * len(pad) + [1] * n
    return torch.tensor([ids], dtype=torch.int32), torch.tensor([m], dtype=torch.int32)
