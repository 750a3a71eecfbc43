"""Model configurations and weight manifests of the Flux.2 denoising path, as the product's own host-side mirror of

  Configuration/Flux2Config.swift:210-329   Flux2TransformerConfig + the dev / klein-4b / klein-9b presets
  Configuration/VAEConfig.swift:7-81        VAEConfig + the small-decoder preset
  Loading/WeightLoader.swift:99-204,397-547 the flattened module keys a checkpoint is handed over under

so that callers (bench.py, the integration shim) need nothing outside this package to create a context, enumerate the
tensors it expects and size synthetic weights. No arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple


@dataclass
class Flux2TransformerConfig:
    """Flux2TransformerConfig (Configuration/Flux2Config.swift:210-289); defaults = Flux.2 Dev (:291-300)."""
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
        return int(float(self.inner_dim) * self.mlp_ratio)  # Int(Float(dim) * mlpRatio), Flux2TransformerBlock.swift:53


def flux2_dev() -> Flux2TransformerConfig:  # Flux2Config.swift:291-300
    return Flux2TransformerConfig()


def klein_4b() -> Flux2TransformerConfig:  # Flux2Config.swift:302-312
    return Flux2TransformerConfig(num_layers=5, num_single_layers=20, num_attention_heads=24, joint_attention_dim=7680,
                                  guidance_embeds=False)


def klein_9b() -> Flux2TransformerConfig:  # Flux2Config.swift:321-329
    return Flux2TransformerConfig(num_layers=8, num_single_layers=24, num_attention_heads=32, joint_attention_dim=12288,
                                  guidance_embeds=False)


PRESETS = {"dev": flux2_dev, "klein4b": klein_4b, "klein9b": klein_9b}


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


def vae_standard() -> VAEConfig:
    return VAEConfig()


def vae_small_decoder() -> VAEConfig:  # VAEConfig.swift:79-81; the pipeline default (Flux2Pipeline.swift:303)
    return VAEConfig(decoder_block_out_channels=(96, 192, 384, 384))


def dit_weight_manifest(cfg: Flux2TransformerConfig, norm_weights: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Flattened Swift module key -> shape of every DiT tensor (Linear weights [out, in], no biases; optionally the
    QK-RMSNorm weights [128]). Key tables: WeightLoader.swift:99-204."""
    D, Hm = cfg.inner_dim, cfg.mlp_hidden
    s: Dict[str, Tuple[int, ...]] = {
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
        for n in ("attn.toQ", "attn.toK", "attn.toV", "attn.addQProj", "attn.addKProj", "attn.addVProj", "attn.toOut",
                  "attn.toAddOut"):
            s[p + n + ".weight"] = (D, D)
        for ff in ("ff", "ffContext"):
            s[p + ff + ".activation.proj.weight"] = (2 * Hm, D)
            s[p + ff + ".linearOut.weight"] = (D, Hm)
        if norm_weights:
            for n in ("normQ", "normK", "normAddedQ", "normAddedK"):
                s[p + f"attn.{n}.weight"] = (cfg.attention_head_dim,)
    for i in range(cfg.num_single_layers):
        p = f"singleTransformerBlocks.{i}."
        s[p + "attn.toQkvMlp.weight"] = (3 * D + 2 * Hm, D)
        s[p + "attn.toOut.weight"] = (D, D + Hm)
        if norm_weights:
            for n in ("normQ", "normK"):
                s[p + f"attn.{n}.weight"] = (cfg.attention_head_dim,)
    return s


def vae_weight_manifest(cfg: VAEConfig, encoder: bool = False) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the VAE decoder (and optionally encoder) tensors: conv weights OHWI (WeightLoader.swift:496-498) + bias,
    GroupNorm weight / bias, the mid-block attention linears with bias, latentBatchNorm running stats over the 128 patchified
    channels (key layout WeightLoader.swift:397-547)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, k, k, ci); s[name + ".bias"] = (co,)

    def lin(name, co, ci):
        s[name + ".weight"] = (co, ci); s[name + ".bias"] = (co,)

    def norm(name, c):
        s[name + ".weight"] = (c,); s[name + ".bias"] = (c,)

    def resnet(name, ci, co):
        norm(name + ".norm1", ci); conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co); conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".convShortcut", co, ci, 1)

    ch, L = cfg.decoder_channels, cfg.latent_channels
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
        for j in range(cfg.layers_per_block + 1):  # VAEDecoder.swift:57
            resnet(f"decoder.upBlocks.{i}.0.{j}", prev if j == 0 else co, co)
        prev = co
        if i < 3:
            conv(f"decoder.upBlocks.{i}.1.conv", co, co, 3)
    norm("decoder.convNormOut", ch[0])
    conv("decoder.convOut", cfg.out_channels, ch[0], 3)
    s["latentBatchNorm.runningMean"] = (4 * L,)
    s["latentBatchNorm.runningVar"] = (4 * L,)
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
    return s


def dit_flops(cfg: Flux2TransformerConfig, S_img: int, S_txt: int = 512) -> Tuple[int, int]:
    """Algorithmic FLOPs of one DiT forward: (GEMM, attention) — the accounting of BASELINE.md §3 / SURVEY §8d."""
    D, Hm = cfg.inner_dim, cfg.mlp_hidden
    S = S_img + S_txt
    g = lambda M, N, K: 2 * M * N * K
    gemm = g(S_img, D, cfg.in_channels) + g(S_txt, D, cfg.joint_attention_dim) + g(S_img, cfg.out_channels, D)
    for s in (S_img, S_txt):
        gemm += cfg.num_layers * (4 * g(s, D, D) + g(s, 2 * Hm, D) + g(s, D, Hm))
    gemm += cfg.num_single_layers * (g(S, 3 * D + 2 * Hm, D) + g(S, D, D + Hm))
    attn = (cfg.num_layers + cfg.num_single_layers) * 4 * S * S * D
    return gemm, attn
