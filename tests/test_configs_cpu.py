"""CPU: the product's own presets / manifests (flux2b.configs) agree with the oracle's restatement of the Swift config types
(Configuration/Flux2Config.swift:291-329, VAEConfig.swift:74-81) — bench.py builds its contexts from the product's tables only."""
import dataclasses


def test_presets_and_manifests_match_oracle():
    from flux2b import configs
    from oracle import flux2_oracle as O
    for mine, ref in ((configs.flux2_dev, O.flux2_dev), (configs.klein_4b, O.klein_4b), (configs.klein_9b, O.klein_9b)):
        a, b = mine(), ref()
        assert dataclasses.asdict(a) == dataclasses.asdict(b)
        assert a.inner_dim == b.inner_dim and a.mlp_hidden == b.mlp_hidden
        assert configs.dit_weight_manifest(a) == {k: tuple(v) for k, v in O.dit_weight_shapes(b).items()}
        assert configs.dit_flops(a, 4096) == configs.dit_flops(b, 4096)
    for mine, ref in ((configs.vae_small_decoder, O.vae_small_decoder), (configs.vae_standard, O.VAEConfig)):
        a, b = mine(), ref()
        assert dataclasses.asdict(a) == dataclasses.asdict(b) and a.decoder_channels == b.decoder_channels
        for enc in (False, True):
            W = O.random_vae_weights(b, seed=1, encoder=enc)
            assert configs.vae_weight_manifest(a, encoder=enc) == {k: tuple(v.shape) for k, v in W.items()}


def test_flop_accounting_matches_baseline_table():
    # Klein 4B @1024^2 28.30 + 6.52 TF, Klein 9B 64.38 + 11.13, Dev S=16896 928.8 + 392.9
    from flux2b import configs
    g, a = configs.dit_flops(configs.klein_4b(), 4096)
    assert abs(g / 1e12 - 28.30) < 0.01 and abs(a / 1e12 - 6.52) < 0.01
    g, a = configs.dit_flops(configs.klein_9b(), 4096)
    assert abs(g / 1e12 - 64.38) < 0.01 and abs(a / 1e12 - 11.13) < 0.01
    g, a = configs.dit_flops(configs.flux2_dev(), 16384)
    assert abs(g / 1e12 - 928.8) < 0.1 and abs(a / 1e12 - 392.9) < 0.1
