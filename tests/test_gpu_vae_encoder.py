"""GPU parity tests of the VAE encoder row through the C ABI: AutoencoderKLFlux2.encode
(VAE/AutoencoderKL.swift:90-127, VAE/VAEEncoder.swift:85-115, asymmetric-pad stride-2 downsample ResnetBlock.swift:189-213)
and encodeImageToPackedSequence / the per-image body of encodeReferenceImages (Flux2Pipeline+ChainHelpers.swift:75-101,
Flux2Pipeline.swift:2196-2213), against the oracle on identical random-init weights.

Tolerance: f16 activations through ~25 convolutions and GroupNorms -> rel-L2 <= 5e-3 (same bar as the decoder)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 5e-3


def _ctx(flux2b, O, vcfg, seed=3):
    VW = O.random_vae_weights(vcfg, seed=seed, encoder=True)
    ctx = flux2b.Context(vae=vcfg)
    ctx.load_weights(VW)
    ctx.finalize()
    return ctx, VW


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (1, 96, 160), (2, 32, 64)])
def test_vae_encode_vs_oracle(flux2b, B, H, W):
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder()          # standard encoder [128, 256, 512, 512] + small decoder (the pipeline default)
    ctx, VW = _ctx(flux2b, O, vcfg)
    img = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(H + W)) * 2 - 1
    lat = ctx.vae_encode(img.numpy())
    assert lat.shape == (B, 32, H // 8, W // 8)
    ref = O.vae_encode(VW, vcfg, img)
    e = rel_l2(lat, ref)
    print(f"vae encode {B}x{H}x{W}: mean rel-L2 {e:.2e}")
    assert e < TOL
    # samplePosterior: mean + exp(logvar / 2) * noise with the caller's noise
    noise = torch.randn(B, 32, H // 8, W // 8, generator=torch.Generator().manual_seed(5))
    lat_s = ctx.vae_encode(img.numpy(), noise.numpy())
    assert rel_l2(lat_s, O.vae_encode(VW, vcfg, img, noise)) < TOL
    # packed, BatchNorm-normalised sequence as the chains consume it
    seq = ctx.encode_image_to_sequence(img.numpy())
    assert seq.shape == (B, (H // 16) * (W // 16), 128)
    want = O.encode_image_to_packed_sequence(VW, vcfg, img)
    assert rel_l2(seq, want) < TOL
    # the plumbing behind the encoder is exact up to the last ulp of the BatchNorm division (the permutes are bit-exact)
    pat = O.normalize_latents_bn(O.pack_latents_to_patchified(torch.from_numpy(np.asarray(lat))),
                                 VW["latentBatchNorm.runningMean"], VW["latentBatchNorm.runningVar"], 1e-4)
    assert np.allclose(np.asarray(seq), O.pack_patchified_to_sequence(pat).numpy(), rtol=1e-6, atol=1e-7)
    ctx.close()


def test_vae_encode_top_left_border_is_not_shifted(flux2b):
    """The downsample pads bottom / right only (ResnetBlock.swift:203-213): a symmetric pad shifts the sampling grid and
    corrupts the top / left latent border — compare the border rows / columns on their own."""
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder()
    ctx, VW = _ctx(flux2b, O, vcfg, seed=8)
    img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1)) * 2 - 1
    lat = torch.from_numpy(np.asarray(ctx.vae_encode(img.numpy())))
    ref = O.vae_encode(VW, vcfg, img)
    for sl in ((slice(None), slice(None), 0), (slice(None), slice(None), slice(None), 0),
               (slice(None), slice(None), -1), (slice(None), slice(None), slice(None), -1)):
        assert rel_l2(lat[sl], ref[sl]) < 2 * TOL
    ctx.close()


def test_vae_decode_of_encode_roundtrip_shapes_and_errors(flux2b):
    from oracle import flux2_oracle as O
    vcfg = O.vae_small_decoder()
    ctx, VW = _ctx(flux2b, O, vcfg)
    img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)) * 2 - 1
    lat = ctx.vae_encode(img.numpy())
    out = ctx.vae_decode(lat)                       # encoder and decoder share the activation pool
    assert out.shape == (1, 3, 64, 64)
    assert rel_l2(out, O.vae_decode(VW, vcfg, O.vae_encode(VW, vcfg, img))) < 2 * TOL
    with pytest.raises(flux2b.Flux2Error) as e:
        ctx.vae_encode(np.zeros((1, 3, 60, 64), dtype=np.float32))
    assert e.value.case == "invalidConfiguration"
    # a context without encoder tensors refuses loudly
    dec_only = flux2b.Context(vae=vcfg)
    dec_only.load_weights(O.random_vae_weights(vcfg, seed=3))
    dec_only.finalize()
    with pytest.raises(flux2b.Flux2Error) as e:
        dec_only.vae_encode(img.numpy())
    assert e.value.case == "modelNotLoaded"
    ctx.close(); dec_only.close()


def test_vae_encoder_golden(flux2b):
    """committed fixture (tests/golden/golden_vae_encoder.npz, tools/make_golden.py encoder)"""
    import os
    from oracle import flux2_oracle as O
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_vae_encoder.npz")))
    vcfg = O.vae_small_decoder()
    ctx, _ = _ctx(flux2b, O, vcfg, seed=3)
    assert rel_l2(ctx.vae_encode(g["img"]), g["moments"][:, :32]) < TOL
    assert rel_l2(ctx.encode_image_to_sequence(g["img"]), g["seq"]) < TOL
    ctx.close()
