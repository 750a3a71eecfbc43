#!/usr/bin/env python
"""Pins the oracle against REAL MLX. Run on a machine that has `mlx` (any Apple-silicon Mac: `pip install mlx`; the reference
links mlx-swift 0.31.6 = MLX core 0.31.x, Package.swift:18):

    python tools/mlx_fixtures.py            # writes tests/golden/mlx_pins.npz
    python -m pytest tests/test_mlx_pins_cpu.py tests -m "not gpu" -q

This container has neither swift nor mlx, so every arithmetic rule that lives inside mlx-swift — the five
quantizers, rms_norm, scaled_dot_product_attention, layer_norm, conv2d, float -> uint8 casts, mixed f16 x f32 promotion — is restated
from the published semantics and the parity claim of this repo is "unpinned" until this file exists. The script uses nothing but
`mlx.core` and numpy; inputs are the committed golden matrix (tests/golden/golden.npz: zero, constant, outlier and all-negative
groups) plus seeded numpy tensors, so the outputs are reproducible bit for bit. tests/test_mlx_pins_cpu.py consumes the file when
present (oracle == MLX: packed codes / scales / biases bit-exact, float ops within accumulation-order noise) and reports
"parity unpinned" as a skip when absent; tests/test_gpu_mlx_pins.py does the same for the CUDA packers on the GPU box.

Call sites being pinned: quantize(model:) Flux2Pipeline.swift:567-578, quantized()/dequantized() WeightLoader.swift:795-815,
MLXFast.rmsNorm Flux2Attention.swift:24, MLXFast.scaledDotProductAttention Flux2Attention.swift:168-174, LayerNorm
Flux2TransformerBlock.swift:56-61, Conv2d ResnetBlock.swift:141-154, asType(.uint8) Flux2Pipeline.swift:2438-2444.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "mlx_pins.npz")

MODES = {  # name -> (mode, bits, group_size)   (Configuration/QuantizationConfig.swift:51-60)
    "qint8": ("affine", 8, 64), "int4": ("affine", 4, 64), "mxfp8": ("mxfp8", 8, 32), "mxfp4": ("mxfp4", 4, 32), "nvfp4": ("nvfp4", 4, 16),
}


def main():
    try:
        import mlx.core as mx
    except ImportError:
        sys.exit("mlx is not installed here: run this script on a Mac with `pip install mlx` and commit tests/golden/mlx_pins.npz")
    out = {"mlx_version": np.array(getattr(mx, "__version__", "unknown"))}
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    w16 = golden["quant_w_f16"]                       # [64, 256] float16
    rng = np.random.default_rng(7)
    extra = (rng.standard_normal((32, 512)) * 0.05).astype(np.float16)   # a second matrix, plain Gaussian
    for tag, w in (("golden", w16), ("gauss", extra)):
        out[f"quant_{tag}_w"] = w
        for name, (mode, bits, group) in MODES.items():
            res = mx.quantize(mx.array(w), group_size=group, bits=bits, mode=mode)
            wq, scales = res[0], res[1]
            biases = res[2] if len(res) > 2 else None
            deq = mx.dequantize(wq, scales, biases, group_size=group, bits=bits, mode=mode) if biases is not None else \
                mx.dequantize(wq, scales, group_size=group, bits=bits, mode=mode)
            out[f"quant_{tag}_{name}_packed"] = np.array(wq)
            out[f"quant_{tag}_{name}_scales"] = np.array(scales) if scales.dtype != mx.bfloat16 else np.array(scales.astype(mx.float32))
            out[f"quant_{tag}_{name}_scales_dtype"] = np.array(str(scales.dtype))
            if biases is not None:
                out[f"quant_{tag}_{name}_biases"] = np.array(biases)
            out[f"quant_{tag}_{name}_dequant"] = np.array(deq.astype(mx.float32))
            # the forward the reference runs: fp32 activations x quantized weight (QuantizedLinear)
            x = mx.array(rng.standard_normal((5, w.shape[1])).astype(np.float32))
            y = mx.quantized_matmul(x, wq, scales, biases, transpose=True, group_size=group, bits=bits, mode=mode) if biases is not None else \
                mx.quantized_matmul(x, wq, scales, transpose=True, group_size=group, bits=bits, mode=mode)
            out[f"qmm_{tag}_{name}_x"] = np.array(x)
            out[f"qmm_{tag}_{name}_y"] = np.array(y.astype(mx.float32))
    # ---- rms_norm / layer_norm (eps 1e-6)
    x = rng.standard_normal((2, 3, 16, 128)).astype(np.float32)
    wn = (1 + 0.1 * rng.standard_normal(128)).astype(np.float32)
    out["rms_x"], out["rms_w"] = x, wn
    out["rms_y"] = np.array(mx.fast.rms_norm(mx.array(x), mx.array(wn), 1e-6))
    out["rms_y_f16"] = np.array(mx.fast.rms_norm(mx.array(x.astype(np.float16)), mx.array(wn.astype(np.float16)), 1e-6).astype(mx.float32))
    xl = rng.standard_normal((4, 384)).astype(np.float32)
    out["ln_x"] = xl
    out["ln_y"] = np.array(mx.fast.layer_norm(mx.array(xl), None, None, 1e-6))
    # ---- scaled dot product attention (scale 128^-0.5, no mask; fp32 and f16 inputs)
    q, k, v = (rng.standard_normal((1, 2, 64, 128)).astype(np.float32) for _ in range(3))
    out["sdpa_q"], out["sdpa_k"], out["sdpa_v"] = q, k, v
    out["sdpa_y"] = np.array(mx.fast.scaled_dot_product_attention(mx.array(q), mx.array(k), mx.array(v), scale=128 ** -0.5))
    out["sdpa_y_f16"] = np.array(mx.fast.scaled_dot_product_attention(mx.array(q.astype(np.float16)), mx.array(k.astype(np.float16)),
                                                                      mx.array(v.astype(np.float16)), scale=128 ** -0.5).astype(mx.float32))
    # ---- additive -inf / -1e9 masks (text encoder: createCausalMask, Qwen3Model.swift:196-231)
    m = np.zeros((64, 64), np.float32); m[np.triu_indices(64, 1)] = -np.inf; m[:, 40:] += -1e9
    out["sdpa_mask"] = m
    out["sdpa_y_masked"] = np.array(mx.fast.scaled_dot_product_attention(mx.array(q), mx.array(k), mx.array(v), scale=128 ** -0.5, mask=mx.array(m)))
    # ---- Linear with f16 weights and fp32 activations (promotion: result dtype and value)
    xw = rng.standard_normal((3, 256)).astype(np.float32)
    wl = (rng.standard_normal((8, 256)) / 16).astype(np.float16)
    y = mx.array(xw) @ mx.array(wl).T
    out["lin_x"], out["lin_w"], out["lin_y"] = xw, wl, np.array(y.astype(mx.float32))
    out["lin_y_dtype"] = np.array(str(y.dtype))
    # ---- conv2d NHWC / OHWI, padding 1; stride 2 with explicit bottom / right padding (ResnetBlock.swift:203-213)
    xc = rng.standard_normal((1, 6, 8, 16)).astype(np.float32)
    wc = (rng.standard_normal((8, 3, 3, 16)) / 12).astype(np.float32)
    out["conv_x"], out["conv_w"] = xc, wc
    out["conv_y"] = np.array(mx.conv2d(mx.array(xc), mx.array(wc), stride=1, padding=1))
    xp = mx.pad(mx.array(xc), [(0, 0), (0, 1), (0, 1), (0, 0)])
    out["conv_y_s2"] = np.array(mx.conv2d(xp, mx.array(wc), stride=2, padding=0))
    # ---- float -> uint8 (postprocessVAEOutput, Flux2Pipeline.swift:2438-2444): truncation or rounding?
    f = np.linspace(-1.2, 1.3, 2001).astype(np.float32)
    out["u8_in"] = f
    out["u8_out"] = np.array(mx.clip((mx.array(f) + 1) * 127.5, 0, 255).astype(mx.uint8))
    # ---- exp / silu references used by SwiGLU
    xs = np.linspace(-12, 12, 4001).astype(np.float32)
    out["silu_x"] = xs
    out["silu_y"] = np.array(mx.array(xs) * mx.sigmoid(mx.array(xs)))
    # ---- the RNG stream the pipeline seeds (LatentUtils.swift:33-41): documented, not reproducible off-MLX
    mx.random.seed(42)
    out["rng_normal_seed42"] = np.array(mx.random.normal((4, 8)))
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, mlx {out['mlx_version']}")


if __name__ == "__main__":
    main()
