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
