"""CPU: the safetensors header reader and the metadata / payload-integrity gate of the pre-quantized checkpoint
(Flux2PrequantizedCheckpoint.isValid, Loading/PrequantizedCheckpoint.swift:107-202) — no device needed."""
import json
import os
import struct

import numpy as np
import pytest


def write_safetensors(path, tensors, metadata, pad_header_to=8, shuffle=False):
    """minimal writer in the layout MLX / safetensors emit: u64 header length, JSON header, raw payload"""
    names = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.uint32): "U32", np.dtype(np.uint8): "U8"}
    hdr, off, blobs = {"__metadata__": metadata}, 0, []
    items = list(tensors.items())
    if shuffle:
        items = items[::-1]
    for k, a in items:
        b = np.ascontiguousarray(a).tobytes()
        hdr[k] = {"dtype": names[a.dtype], "shape": list(a.shape), "data_offsets": [off, off + len(b)]}
        off += len(b)
        blobs.append(b)
    h = json.dumps(hdr, indent=1 if shuffle else None).encode()
    h += b" " * (-len(h) % pad_header_to)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)) + h + b"".join(blobs))


def meta(quant="nvfp4", bits=4, group=16, mode="nvfp4", **over):
    m = {"format": "flux2-mlx-prequantized-v1", "quantization": quant, "bits": str(bits), "group_size": str(group), "mode": mode,
         "component": "transformer", "source": "FLUX.2-klein-4B", "source_fingerprint": "a.safetensors:10:1700000000",
         "created_by": "flux-2-swift-mlx"}
    m.update(over)
    return m


def test_is_valid_gate(tmp_path):
    import flux2b
    T = {"xEmbedder.weight": np.arange(64, dtype=np.uint32).reshape(4, 16), "xEmbedder.scales": np.full((4, 8), 0x38, np.uint8),
         "transformerBlocks.0.attn.normQ.weight": np.ones(128, np.float16)}
    p = str(tmp_path / "transformer.safetensors")
    write_safetensors(p, T, meta())
    assert flux2b.prequantized_is_valid(p, "nvfp4", "FLUX.2-klein-4B", "a.safetensors:10:1700000000")
    assert flux2b.prequantized_is_valid(p, "nvfp4")                          # source / fingerprint not checked when not given
    assert not flux2b.prequantized_is_valid(p, "mxfp4") and "quantization" in flux2b.last_error()
    assert not flux2b.prequantized_is_valid(p, "nvfp4", "FLUX.2-klein-9B") and "source" in flux2b.last_error()
    assert not flux2b.prequantized_is_valid(p, "nvfp4", "FLUX.2-klein-4B", "a.safetensors:11:1700000000") and "stale" in flux2b.last_error()
    assert not flux2b.prequantized_is_valid(str(tmp_path / "absent.safetensors"), "nvfp4")
    # every metadata field of PrequantizedCheckpoint.swift:177-185 is required
    for k, v in (("format", "flux2-mlx-prequantized-v0"), ("bits", "8"), ("group_size", "32"), ("mode", "affine"), ("component", "vae")):
        write_safetensors(p, T, meta(**{k: v}))
        assert not flux2b.prequantized_is_valid(p, "nvfp4"), k
    # pretty-printed header, reversed key order: still the same file
    write_safetensors(p, T, meta(), shuffle=True)
    assert flux2b.prequantized_is_valid(p, "nvfp4")


def test_truncated_or_padded_payload_is_rejected(tmp_path):
    """a valid header over truncated data must not pass (PrequantizedCheckpoint.swift:99-141)"""
    import flux2b
    T = {"xEmbedder.weight": np.arange(256, dtype=np.uint32).reshape(16, 16), "xEmbedder.scales": np.zeros((16, 8), np.uint8)}
    p = str(tmp_path / "t.safetensors")
    write_safetensors(p, T, meta())
    assert flux2b.prequantized_is_valid(p, "nvfp4")
    blob = open(p, "rb").read()
    for cut in (1, 100, len(blob) - 9):
        open(p, "wb").write(blob[:-cut])
        assert not flux2b.prequantized_is_valid(p, "nvfp4"), cut
    open(p, "wb").write(blob + b"\0")
    assert not flux2b.prequantized_is_valid(p, "nvfp4")
    open(p, "wb").write(blob[:4])
    assert not flux2b.prequantized_is_valid(p, "nvfp4")
    open(p, "wb").write(struct.pack("<Q", 1 << 40) + b"{}")
    assert not flux2b.prequantized_is_valid(p, "nvfp4")
    open(p, "wb").write(struct.pack("<Q", 8) + b"{notjson")
    assert not flux2b.prequantized_is_valid(p, "nvfp4")


def test_file_written_by_the_official_safetensors_library_is_read(tmp_path):
    """reader pinned against the reference implementation of the format: a file produced by `safetensors.numpy.save_file`
    (its own key order, header padding and alignment) passes the metadata / payload gate, and a metadata edit fails it"""
    st = pytest.importorskip("safetensors.numpy")
    import flux2b
    T = {"xEmbedder.weight": np.arange(64, dtype=np.uint32).reshape(4, 16), "xEmbedder.scales": np.full((4, 8), 0x38, np.uint8),
         "transformerBlocks.0.attn.normQ.weight": np.ones(128, np.float16), "singleTransformerBlocks.3.attn.normK.weight": np.ones(128, np.float32)}
    p = str(tmp_path / "transformer.safetensors")
    st.save_file(T, p, metadata=meta())
    assert flux2b.prequantized_is_valid(p, "nvfp4", "FLUX.2-klein-4B", "a.safetensors:10:1700000000"), flux2b.last_error()
    st.save_file(T, p, metadata=meta(bits="8"))
    assert not flux2b.prequantized_is_valid(p, "nvfp4")
