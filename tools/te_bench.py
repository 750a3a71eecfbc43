"""Times the text-embedding producer on one B200: Klein's Qwen3 encoders at their real shapes, 512 tokens,
hidden states of layers 9 / 18 / 27 (only 27 layers are built and run). Random-init bf16 weights generated on the device.
usage: python tools/te_bench.py [qwen3_4b qwen3_8b] -> one JSON line per model on stdout."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "flux-2-swift-mlx_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch

import flux2b
from oracle import flux2_oracle as O   # configs / weight shapes only (bench-side, never on the product path)


def run(name, reps=10, profile_one=False):
    cfg = getattr(O, name)()
    layers = O.KLEIN_HIDDEN_STATE_LAYERS
    te = flux2b.TextEncoder(cfg, options={"keep_raw_weights": 0})
    g = torch.Generator(device="cuda").manual_seed(0)
    for k, shp in O.te_weight_shapes(cfg, layers=max(layers)).items():
        if len(shp) == 1:
            t = torch.ones(shp, device="cuda")
        elif k == "model.embed_tokens.weight":
            t = (torch.randn(shp, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        else:
            t = ((torch.rand(shp, device="cuda", generator=g) * 2 - 1) / shp[1] ** 0.5).to(torch.bfloat16)
        te.set_tensor(k, t)
    te.finalize()
    torch.cuda.empty_cache()
    ex = flux2b.KleinEmbeddingExtractor(te)
    toks = np.random.default_rng(0).integers(0, cfg.vocab_size, 77).tolist()
    for _ in range(3):
        out = ex.extract(toks)
    te.synchronize()
    if profile_one:   # for `ncu --profile-from-start off`: one prefill between cudaProfilerStart / Stop, no bench line
        torch.cuda.synchronize(); torch.cuda.profiler.start()
        ex.extract(toks)
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
        te.close()
        return
    t0 = time.perf_counter()
    for _ in range(reps):
        out = ex.extract(toks)              # host ids in, host fp32 [1, 512, 3 * hidden] out: end to end
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    # the same with the result left on the device in bf16 (what a pipeline that feeds the DiT directly does): no 15 - 25 MB D2H
    ids_np = np.asarray([toks + [ex.PAD_TOKEN_ID] * (512 - len(toks))], dtype=np.int32)
    mask_np = np.asarray([[1] * len(toks) + [0] * (512 - len(toks))], dtype=np.int32)
    out_dev = torch.empty((1, 512, 3 * cfg.hidden_size), dtype=torch.bfloat16, device="cuda")
    li = (flux2b.ctypes.c_int * 3)(*layers)
    def dev_call():
        flux2b._ck(flux2b.lib().flux2b_te_hidden_states(te._h, 1, 512, flux2b._ptr(ids_np), flux2b._ptr(mask_np), li, 3,
                                                        flux2b._ptr(out_dev), flux2b.BF16))
    for _ in range(3):
        dev_call()
    te.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev_call()
    te.synchronize()
    dev_ms = (time.perf_counter() - t0) / reps * 1e3
    te.set_option("te_graph", 0)
    for _ in range(2):
        dev_call()
    te.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev_call()
    te.synchronize()
    dev_nograph_ms = (time.perf_counter() - t0) / reps * 1e3
    te.set_option("te_graph", 1)
    te.prof_enable(True); te.prof_reset()
    l0 = te.launch_count()
    ex.extract(toks)
    prof = {n: te.prof_get(k) for k, n in enumerate(("gemm", "attn", "elem"))}
    launches = te.launch_count() - l0
    kern_ms = sum(p["ms"] for p in prof.values())
    Hd, I, Nq, Nkv = cfg.hidden_size, cfg.intermediate_size, cfg.num_heads * 128, cfg.num_kv_heads * 128
    flops = max(layers) * 2.0 * 512 * (Hd * (Nq + 2 * Nkv) + Nq * Hd + 3 * Hd * I)
    line = {"what": "text_embedding_producer", "model": name, "tokens": 512, "layers_run": max(layers), "hidden_state_layers": list(layers),
            "e2e_ms": e2e_ms, "device_out_bf16_ms": dev_ms, "device_out_bf16_no_graph_ms": dev_nograph_ms, "kernel_ms": kern_ms, "launches": launches, "gemm_tflops": prof["gemm"]["flops"] / prof["gemm"]["ms"] / 1e9,
            "gemm_ms": prof["gemm"]["ms"], "attn_ms": prof["attn"]["ms"], "elem_ms": prof["elem"]["ms"],
            "gemm_flops_check": flops / prof["gemm"]["flops"], "out_shape": list(out.shape), "finite": bool(np.isfinite(out).all())}
    print(json.dumps(line), flush=True)
    te.close()


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for n in (args or ["qwen3_4b", "qwen3_8b"]):
        run(n, profile_one="--profile-one" in sys.argv)
