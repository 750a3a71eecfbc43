"""ctypes front-end of oracle/quant_oracle.c — TEST INFRASTRUCTURE ONLY (see the header of that file)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

QUANT = {"bf16": 0, "qint8": 1, "int4": 2, "mxfp8": 3, "mxfp4": 4, "nvfp4": 5}


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return os.path.join(_HERE, "_build", "libquant_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libquant_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_quantize.restype = ctypes.c_int
        L.oracle_dequantize.restype = ctypes.c_int
        L.oracle_to_e8m0.argtypes = [ctypes.c_float]; L.oracle_to_e8m0.restype = ctypes.c_uint8
        L.oracle_to_e4m3.argtypes = [ctypes.c_float]; L.oracle_to_e4m3.restype = ctypes.c_uint8
        L.oracle_to_e2m1.argtypes = [ctypes.c_float]; L.oracle_to_e2m1.restype = ctypes.c_uint8
        L.oracle_from_e8m0.argtypes = [ctypes.c_uint8]; L.oracle_from_e8m0.restype = ctypes.c_float
        L.oracle_from_e4m3.argtypes = [ctypes.c_uint8]; L.oracle_from_e4m3.restype = ctypes.c_float
        L.oracle_from_e2m1.argtypes = [ctypes.c_uint8]; L.oracle_from_e2m1.restype = ctypes.c_float
        _LIB = L
    return _LIB


def params(quant: int):
    b, g, h = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    if lib().oracle_quant_params(quant, ctypes.byref(b), ctypes.byref(g), ctypes.byref(h)) != 0:
        raise ValueError("not a quantized mode")
    return b.value, g.value, bool(h.value)


def _dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return 0
    if a.dtype == np.float16:
        return 1
    if a.dtype == np.uint16:  # raw bf16 bits
        return 2
    raise TypeError(a.dtype)


def quantize(quant: int, w: np.ndarray):
    """w [rows, cols] float32 | float16 | uint16(bf16 bits) -> (packed uint32 [rows, cols*bits/32], scales, biases|None).
    scales: float16 for affine modes, uint8 otherwise (MLX layout)."""
    bits, group, has_b = params(quant)
    w = np.ascontiguousarray(w)
    rows, cols = w.shape
    packed = np.zeros((rows, cols * bits // 32), dtype=np.uint32)
    scales = np.zeros((rows, cols // group), dtype=np.float16 if has_b else np.uint8)
    biases = np.zeros((rows, cols // group), dtype=np.float16)
    rc = lib().oracle_quantize(quant, w.ctypes.data_as(ctypes.c_void_p), _dtype_code(w), ctypes.c_int64(rows),
                               ctypes.c_int64(cols), packed.ctypes.data_as(ctypes.c_void_p),
                               scales.ctypes.data_as(ctypes.c_void_p), biases.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("oracle_quantize failed")
    return packed, scales, (biases if has_b else None)


def dequantize(quant: int, packed: np.ndarray, scales: np.ndarray, biases, cols: int) -> np.ndarray:
    bits, group, has_b = params(quant)
    rows = packed.shape[0]
    out = np.zeros((rows, cols), dtype=np.float32)
    packed = np.ascontiguousarray(packed)
    scales = np.ascontiguousarray(scales)
    b = np.ascontiguousarray(biases) if has_b else np.zeros(1, dtype=np.float16)
    rc = lib().oracle_dequantize(quant, packed.ctypes.data_as(ctypes.c_void_p), scales.ctypes.data_as(ctypes.c_void_p),
                                 b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(rows), ctypes.c_int64(cols),
                                 out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("oracle_dequantize failed")
    return out


def fake_quant_activation(quant: int, x, f16: bool = False):
    """Checker model of the product's on-the-fly activation quantisation in native block-scaled mode (csrc/quant.cu:
    mx_quantize_act): x (torch fp32 [..., K]) -> rounded to the 16-bit operand type -> quantised -> dequantised fp32.
    mxfp4 / nvfp4 use exactly the weight packer above (groups along the last dim); mxfp8 uses E4M3 elements with the
    scale 2^ceil(log2(amax / 448)) per 32 elements (nothing saturates)."""
    import torch
    shp = x.shape
    K = shp[-1]
    x16 = x.to(torch.float16 if f16 else torch.bfloat16).reshape(-1, K)
    if quant in (4, 5):
        raw = x16.view(torch.int16).numpy().view(np.uint16) if not f16 else x16.numpy()
        p, s, _ = quantize(quant, raw)
        return torch.from_numpy(dequantize(quant, p, s, None, K)).reshape(shp)
    if quant != 3:
        raise ValueError("native block-scaled modes are mxfp8, mxfp4, nvfp4")
    g = x16.float().reshape(-1, K // 32, 32)
    amax = g.abs().amax(dim=-1, keepdim=True)
    qv = (amax * np.float32(1.0 / 448.0)).contiguous()
    u = qv.view(torch.int32)
    e = ((u >> 23) & 0xff) - 127 + ((u & 0x7fffff) != 0).to(torch.int32)
    e = torch.where(amax > 0, e.clamp(-127, 127), torch.full_like(e, -127))
    inv = torch.ldexp(torch.ones_like(amax), -e)
    q8 = (g * inv).to(torch.float8_e4m3fn).float()
    return (q8 * torch.ldexp(torch.ones_like(amax), e)).reshape(shp)
