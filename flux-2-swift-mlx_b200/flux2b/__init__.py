"""flux2b — Python (ctypes) front-end of libflux2b.so, the C-ABI library of include/flux2b.h.

This is the harness the parity tests and bench.py drive; it mirrors the reference's Swift surface for the hot path
(same names, argument meaning and error behaviour) so tests read like the reference's own:

    Flux2Transformer2DModel.__call__      <- Transformer/Flux2Transformer.swift:123
    FlowMatchEulerScheduler               <- Scheduler/FlowMatchEulerScheduler.swift:34
    AutoencoderKLFlux2.decode             <- VAE/AutoencoderKL.swift:129
    LatentUtils.*                         <- Pipeline/LatentUtils.swift
    Flux2Pipeline.denoise / generate      <- Pipeline/Flux2Pipeline.swift:1933-2098 (loop body + tail)

There is no Python or CPU compute here: every call crosses the C ABI; arguments may be numpy arrays (host) or torch
tensors (host or CUDA). Importing works without a GPU (symbol checks); creating a context raises Flux2Error.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLUX2B_LIB: A/B aid — load another build of the same library (kernel variants measured on the same GPU box)
LIB_PATH = os.environ.get("FLUX2B_LIB") or os.path.join(os.path.dirname(_HERE), "libflux2b.so")

F32, F16, BF16, U32, U8, I32 = 0, 1, 2, 3, 4, 5
QUANT = {"bf16": 0, "qint8": 1, "int4": 2, "mxfp8": 3, "mxfp4": 4, "nvfp4": 5}
PROF_GEMM, PROF_ATTN, PROF_ELEMWISE, PROF_CONV, PROF_GEMV, PROF_COMM, PROF_GROUPNORM = range(7)

_STATUS = {-1: "modelNotLoaded", -2: "invalidConfiguration", -3: "insufficientMemory", -4: "weightLoadingFailed",
           -5: "imageProcessingFailed", -6: "generationFailed", -7: "generationCancelled", -8: "noDevice", -9: "cuda"}


class Flux2Error(RuntimeError):
    """Mirror of Flux2Error (Flux2Core.swift:14-40)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"Flux2Error.{_STATUS.get(code, code)}: {message}")
        self.code = code
        self.case = _STATUS.get(code, str(code))


class DitConfigC(ctypes.Structure):
    _fields_ = [("patch_size", ctypes.c_int), ("in_channels", ctypes.c_int), ("out_channels", ctypes.c_int),
                ("num_layers", ctypes.c_int), ("num_single_layers", ctypes.c_int),
                ("attention_head_dim", ctypes.c_int), ("num_attention_heads", ctypes.c_int),
                ("joint_attention_dim", ctypes.c_int), ("guidance_embeds", ctypes.c_int),
                ("axes_dims_rope", ctypes.c_int * 4), ("rope_theta", ctypes.c_float), ("mlp_ratio", ctypes.c_float)]


class VaeConfigC(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int), ("out_channels", ctypes.c_int), ("latent_channels", ctypes.c_int),
                ("layers_per_block", ctypes.c_int), ("norm_num_groups", ctypes.c_int),
                ("decoder_channels", ctypes.c_int * 4), ("norm_eps", ctypes.c_float), ("encoder_channels", ctypes.c_int * 4)]


class TeConfigC(ctypes.Structure):
    _fields_ = [("vocab_size", ctypes.c_int), ("hidden_size", ctypes.c_int), ("intermediate_size", ctypes.c_int),
                ("num_layers", ctypes.c_int), ("num_heads", ctypes.c_int), ("num_kv_heads", ctypes.c_int),
                ("head_dim", ctypes.c_int), ("qk_norm", ctypes.c_int), ("rms_norm_eps", ctypes.c_float),
                ("rope_theta", ctypes.c_float), ("max_position_embeddings", ctypes.c_int)]


class StepContextC(ctypes.Structure):
    _fields_ = [("step_idx", ctypes.c_int), ("total_steps", ctypes.c_int), ("sigma", ctypes.c_float),
                ("sigma_next", ctypes.c_float), ("height", ctypes.c_int), ("width", ctypes.c_int), ("is_i2i", ctypes.c_int)]


HOOK_T = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(StepContextC), ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                          ctypes.c_void_p)


class DenoiseParamsC(ctypes.Structure):
    _fields_ = [("height", ctypes.c_int), ("width", ctypes.c_int), ("num_sigmas", ctypes.c_int),
                ("sigmas", ctypes.POINTER(ctypes.c_float)), ("guidance", ctypes.c_void_p), ("cfg_scale", ctypes.c_float),
                ("enc", ctypes.c_void_p), ("enc_uncond", ctypes.c_void_p), ("enc_dtype", ctypes.c_int),
                ("S_txt", ctypes.c_int), ("ref_latents", ctypes.c_void_p), ("ref_ids", ctypes.c_void_p),
                ("S_ref", ctypes.c_int), ("hook", HOOK_T), ("hook_user", ctypes.c_void_p), ("kv_cache", ctypes.c_int),
                ("S_txt_uncond", ctypes.c_int)]


EXPORTS = [
    "flux2b_version", "flux2b_last_error", "flux2b_device_count", "flux2b_create", "flux2b_destroy", "flux2b_set_stream",
    "flux2b_synchronize", "flux2b_set_option", "flux2b_set_tensor", "flux2b_get_tensor", "flux2b_finalize_weights",
    "flux2b_merge_lora", "flux2b_quant_params", "flux2b_quantize_matrix", "flux2b_dequantize_matrix", "flux2b_dit_forward",
    "flux2b_dit_forward_kv_extract", "flux2b_dit_forward_kv_cached", "flux2b_kv_cache_clear", "flux2b_get_block_output",
    "flux2b_compute_empirical_mu", "flux2b_scheduler_set_timesteps", "flux2b_scheduler_set_custom_sigmas",
    "flux2b_euler_step", "flux2b_scale_noise", "flux2b_pack_patchified_to_sequence", "flux2b_unpack_sequence_to_patchified",
    "flux2b_unpatchify_latents", "flux2b_pack_latents_to_patchified", "flux2b_bn_latents", "flux2b_image_position_ids",
    "flux2b_text_position_ids", "flux2b_reference_position_ids", "flux2b_vae_decode", "flux2b_vae_decode_u8", "flux2b_vae_encode", "flux2b_encode_image_to_sequence", "flux2b_load_safetensors", "flux2b_save_prequantized",
    "flux2b_load_prequantized", "flux2b_prequantized_is_valid",
    "flux2b_denoise", "flux2b_generate", "flux2b_repaint_blend", "flux2b_sp_unique_id", "flux2b_sp_init", "flux2b_sp_layout",
    "flux2b_prof_enable", "flux2b_prof_reset", "flux2b_prof_get", "flux2b_launch_count", "flux2b_op_gemm", "flux2b_op_gemm_mx", "flux2b_op_gemm_mxfp8", "flux2b_op_linear_quantized",
    "flux2b_op_attention", "flux2b_op_ln_modulate", "flux2b_op_qk_norm_rope", "flux2b_op_rope_table",
    "flux2b_op_timestep_embedding", "flux2b_op_conv2d", "flux2b_op_groupnorm_silu",
    "flux2b_te_create", "flux2b_te_hidden_states", "flux2b_op_attention_causal",
]

_lib = None


def lib():
    """Load libflux2b.so; fails loudly when the CUDA extension is missing (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make -C flux-2-swift-mlx_b200/csrc` "
                              "(or __graft_entry__.build()); flux2b has no CPU / PyTorch fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.flux2b_version.restype = ctypes.c_char_p
        L.flux2b_last_error.restype = ctypes.c_char_p
        L.flux2b_get_tensor.restype = ctypes.c_int64
        L.flux2b_get_block_output.restype = ctypes.c_int64
        L.flux2b_launch_count.restype = ctypes.c_int64
        L.flux2b_compute_empirical_mu.restype = ctypes.c_float
        L.flux2b_compute_empirical_mu.argtypes = [ctypes.c_int, ctypes.c_int]
        L.flux2b_destroy.restype = None
        _lib = L
    return _lib


def _err(code: int):
    raise Flux2Error(code, lib().flux2b_last_error().decode())


def _ck(code: int) -> int:
    if code < 0:
        _err(code)
    return code


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> Optional[ctypes.c_void_p]:
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous(), "tensor must be contiguous"
        return ctypes.c_void_p(x.data_ptr())
    assert x.flags["C_CONTIGUOUS"], "array must be contiguous"
    return ctypes.c_void_p(x.ctypes.data)


def _dtype_code(x) -> int:
    if _is_torch(x):
        import torch
        return {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16, torch.int32: I32, torch.uint8: U8,
                torch.uint32: U32}[x.dtype]
    return {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.uint32): U32, np.dtype(np.uint8): U8,
            np.dtype(np.int32): I32}[x.dtype]


def device_count() -> int:
    return lib().flux2b_device_count()


def last_error() -> str:
    return (lib().flux2b_last_error() or b"").decode()


def prequantized_is_valid(path: str, quant, source_name: Optional[str] = None, source_fingerprint: Optional[str] = None) -> bool:
    """Flux2PrequantizedCheckpoint.isValid (PrequantizedCheckpoint.swift:150-166); runs without a GPU."""
    q = QUANT[quant] if isinstance(quant, str) else int(quant)
    return bool(lib().flux2b_prequantized_is_valid(path.encode(), q, source_name.encode() if source_name is not None else None,
                                                   source_fingerprint.encode() if source_fingerprint is not None else None))


# ------------------------------------------------------------------------------------------------ host-only helpers
def compute_empirical_mu(image_seq_len: int, num_steps: int) -> float:
    """computeEmpiricalMu (FlowMatchEulerScheduler.swift:9-28)."""
    return float(lib().flux2b_compute_empirical_mu(image_seq_len, num_steps))


class FlowMatchEulerScheduler:
    """Mirror of FlowMatchEulerScheduler (Scheduler/FlowMatchEulerScheduler.swift:34-260); the math runs in the library."""

    def __init__(self, num_train_timesteps: int = 1000):
        self.num_train_timesteps = num_train_timesteps
        self.sigmas: List[float] = []
        self.timesteps: List[float] = []
        self.step_index = 0

    def set_timesteps(self, num_inference_steps: int, image_seq_len: Optional[int] = None, strength: float = 1.0) -> int:
        buf = (ctypes.c_float * (num_inference_steps + 1))()
        t0 = ctypes.c_int(0)
        n = _ck(lib().flux2b_scheduler_set_timesteps(num_inference_steps, image_seq_len if image_seq_len else -1,
                                                     ctypes.c_float(strength), buf, ctypes.byref(t0)))
        self.sigmas = [float(buf[i]) for i in range(n)]
        self.timesteps = [s * self.num_train_timesteps for s in self.sigmas]
        self.step_index = 0
        return t0.value

    def set_custom_sigmas(self, sigmas: Sequence[float]) -> None:
        if not sigmas:
            return
        src = (ctypes.c_float * len(sigmas))(*sigmas)
        dst = (ctypes.c_float * (len(sigmas) + 1))()
        n = lib().flux2b_scheduler_set_custom_sigmas(src, len(sigmas), dst)
        self.sigmas = [float(dst[i]) for i in range(n)]
        self.timesteps = [s * self.num_train_timesteps for s in self.sigmas]
        self.step_index = 0

    @property
    def initial_sigma(self) -> float:
        return self.sigmas[0] if self.sigmas else 1.0

    def step(self, ctx: "Context", model_output, sample):
        """In place on `sample` (host or device buffer); returns it."""
        if self.step_index >= len(self.sigmas) - 1:
            return sample
        ctx.euler_step(sample, model_output, self.sigmas[self.step_index], self.sigmas[self.step_index + 1])
        self.step_index += 1
        return sample


def image_position_ids(height: int, width: int) -> np.ndarray:
    n = (height // 16) * (width // 16)
    out = np.zeros((n, 4), dtype=np.int32)
    lib().flux2b_image_position_ids(height, width, _ptr(out))
    return out


def text_position_ids(length: int) -> np.ndarray:
    out = np.zeros((length, 4), dtype=np.int32)
    lib().flux2b_text_position_ids(length, _ptr(out))
    return out


def reference_position_ids(lat_h: Sequence[int], lat_w: Sequence[int], scale: int = 10) -> np.ndarray:
    n = sum(h * w for h, w in zip(lat_h, lat_w))
    out = np.zeros((n, 4), dtype=np.int32)
    hh = (ctypes.c_int * len(lat_h))(*lat_h)
    ww = (ctypes.c_int * len(lat_w))(*lat_w)
    lib().flux2b_reference_position_ids(hh, ww, len(lat_h), scale, _ptr(out))
    return out


class SpLayoutC(ctypes.Structure):
    _fields_ = [("txt_row0", ctypes.c_int), ("txt_rows", ctypes.c_int), ("img_row0", ctypes.c_int), ("img_rows", ctypes.c_int),
                ("local_rows", ctypes.c_int), ("heads_per_rank", ctypes.c_int), ("qkv_chunk_elems", ctypes.c_int64),
                ("o_chunk_elems", ctypes.c_int64)]


def sp_layout(world: int, rank: int, S_txt: int, S_img: int, num_heads: int) -> SpLayoutC:
    """How the joint sequence is sharded for Ulysses sequence parallelism (host-only)."""
    out = SpLayoutC()
    _ck(lib().flux2b_sp_layout(world, rank, S_txt, S_img, num_heads, ctypes.byref(out)))
    return out


def sp_unique_id() -> bytes:
    """128-byte NCCL unique id; create on rank 0 and broadcast (torch.distributed / MPI / a file)."""
    buf = ctypes.create_string_buffer(128)
    _ck(lib().flux2b_sp_unique_id(buf))
    return buf.raw


def quant_params(quant: int):
    b, g, h, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _ck(lib().flux2b_quant_params(quant, ctypes.byref(b), ctypes.byref(g), ctypes.byref(h), ctypes.byref(s)))
    return b.value, g.value, bool(h.value), s.value


# ------------------------------------------------------------------------------------------------ context
class Context:
    """One flux2b_ctx: one GPU, one stream, one set of weights (== one Flux2Pipeline's transformer + VAE)."""

    def __init__(self, dit=None, vae=None, quant: int = 0, device: int = 0, options: Optional[Dict[str, int]] = None):
        L = lib()
        self._h = ctypes.c_void_p()
        self.dit_cfg, self.vae_cfg = dit, vae
        dc = vc = None
        if dit is not None:
            dc = DitConfigC(dit.patch_size, dit.in_channels, dit.out_channels, dit.num_layers, dit.num_single_layers,
                            dit.attention_head_dim, dit.num_attention_heads, dit.joint_attention_dim,
                            int(dit.guidance_embeds), (ctypes.c_int * 4)(*dit.axes_dims_rope), dit.rope_theta, dit.mlp_ratio)
        if vae is not None:
            vc = VaeConfigC(vae.in_channels, vae.out_channels, vae.latent_channels, vae.layers_per_block,
                            vae.norm_num_groups, (ctypes.c_int * 4)(*vae.decoder_channels), vae.norm_eps,
                            (ctypes.c_int * 4)(*getattr(vae, "block_out_channels", (0, 0, 0, 0))))
        _ck(L.flux2b_create(device, ctypes.byref(dc) if dc else None, ctypes.byref(vc) if vc else None, quant,
                            ctypes.byref(self._h)))
        for k, v in (options or {}).items():
            self.set_option(k, v)
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().flux2b_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lifecycle / options
    def set_option(self, name: str, value: int):
        _ck(lib().flux2b_set_option(self._h, name.encode(), int(value)))

    # -- safetensors / pre-quantized checkpoint (Loading/PrequantizedCheckpoint.swift)
    def load_safetensors(self, path: str) -> int:
        n = lib().flux2b_load_safetensors(self._h, path.encode())
        if n < 0:
            _ck(n)
        return n

    def save_prequantized(self, path: str, source_name: str = "", source_fingerprint: str = "unknown", lora_baked: bool = False):
        """Flux2PrequantizedCheckpoint.save(model:sourceModelPath:quantization:component:loRABaked:)"""
        _ck(lib().flux2b_save_prequantized(self._h, path.encode(), source_name.encode(), source_fingerprint.encode(), int(lora_baked)))

    def load_prequantized(self, path: str, source_name: Optional[str] = None, source_fingerprint: Optional[str] = None) -> bool:
        """Flux2PrequantizedCheckpoint.load(into:...): True = applied (finalize next); False = context untouched, fall back
        to the standard load (reason: flux2b.last_error())."""
        r = lib().flux2b_load_prequantized(self._h, path.encode(), source_name.encode() if source_name is not None else None,
                                           source_fingerprint.encode() if source_fingerprint is not None else None)
        if r < 0:
            _ck(r)
        return r == 0

    def set_stream(self, cuda_stream: int):
        _ck(lib().flux2b_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        _ck(lib().flux2b_synchronize(self._h))

    def sp_init(self, unique_id: bytes, rank: int, world: int):
        """Join an Ulysses sequence-parallel group (one context per GPU / process)."""
        assert len(unique_id) == 128
        _ck(lib().flux2b_sp_init(self._h, ctypes.c_char_p(unique_id), rank, world))

    # -- weights
    def set_tensor(self, key: str, t):
        shape = (ctypes.c_int64 * len(t.shape))(*[int(s) for s in t.shape])
        _ck(lib().flux2b_set_tensor(self._h, key.encode(), _ptr(t), _dtype_code(t), shape, len(t.shape)))

    def load_weights(self, weights: Dict[str, object], dtype=None):
        """weights: flattened-Swift-key -> torch tensor; optionally cast floating weights (e.g. torch.bfloat16)."""
        for k, t in weights.items():
            if dtype is not None and _is_torch(t) and t.is_floating_point() and t.dim() >= 2:
                t = t.to(dtype)
            self.set_tensor(k, t.contiguous() if _is_torch(t) else np.ascontiguousarray(t))

    def finalize(self):
        _ck(lib().flux2b_finalize_weights(self._h))

    def get_tensor(self, key: str) -> np.ndarray:
        dt, nd = ctypes.c_int(), ctypes.c_int()
        shape = (ctypes.c_int64 * 6)()
        nbytes = _ck(lib().flux2b_get_tensor(self._h, key.encode(), None, ctypes.c_size_t(0), ctypes.byref(dt), shape, ctypes.byref(nd)))
        npdt = {F32: np.float32, F16: np.float16, BF16: np.uint16, U32: np.uint32, U8: np.uint8, I32: np.int32}[dt.value]
        out = np.zeros([shape[i] for i in range(nd.value)], dtype=npdt)
        _ck(lib().flux2b_get_tensor(self._h, key.encode(), _ptr(out), ctypes.c_size_t(nbytes), None, None, None))
        return out

    def merge_lora(self, layer_path: str, A, B, scale: float):
        _ck(lib().flux2b_merge_lora(self._h, layer_path.encode(), _ptr(A), _ptr(B), int(A.shape[0]), _dtype_code(A),
                                    ctypes.c_float(scale)))

    def quantize_matrix(self, quant: int, w):
        bits, group, has_b, sdt = quant_params(quant)
        rows, cols = w.shape
        packed = np.zeros((rows, cols * bits // 32), dtype=np.uint32)
        scales = np.zeros((rows, cols // group), dtype=np.float16 if has_b else np.uint8)
        biases = np.zeros((rows, cols // group), dtype=np.float16) if has_b else None
        _ck(lib().flux2b_quantize_matrix(self._h, quant, _ptr(w), _dtype_code(w), ctypes.c_int64(rows), ctypes.c_int64(cols),
                                         _ptr(packed), _ptr(scales), _ptr(biases)))
        return packed, scales, biases

    def dequantize_matrix(self, quant: int, packed, scales, biases, cols: int) -> np.ndarray:
        out = np.zeros((packed.shape[0], cols), dtype=np.float32)
        _ck(lib().flux2b_dequantize_matrix(self._h, quant, _ptr(packed), _ptr(scales), _ptr(biases),
                                           ctypes.c_int64(packed.shape[0]), ctypes.c_int64(cols), _ptr(out), F32))
        return out

    # -- DiT
    def dit_forward(self, hidden, enc, timestep, guidance, img_ids, txt_ids, out=None):
        B, S_img, _ = hidden.shape
        S_txt = enc.shape[1]
        if out is None:
            out = _empty_like_backend(hidden, (B, S_img, self.dit_cfg.out_channels))
        _ck(lib().flux2b_dit_forward(self._h, B, S_img, S_txt, _ptr(hidden), _ptr(enc), _dtype_code(enc), _ptr(timestep),
                                     _ptr(guidance), _ptr(img_ids), _ptr(txt_ids), _ptr(out)))
        return out

    def dit_forward_kv_extract(self, hidden, ref_hidden, enc, timestep, guidance, img_ids, ref_ids, txt_ids):
        B, S_img, _ = hidden.shape
        out = _empty_like_backend(hidden, (B, S_img, self.dit_cfg.out_channels))
        _ck(lib().flux2b_dit_forward_kv_extract(self._h, B, S_img, ref_hidden.shape[1], enc.shape[1], _ptr(hidden),
                                                _ptr(ref_hidden), _ptr(enc), _dtype_code(enc), _ptr(timestep), _ptr(guidance),
                                                _ptr(img_ids), _ptr(ref_ids), _ptr(txt_ids), _ptr(out)))
        return out

    def dit_forward_kv_cached(self, hidden, enc, timestep, guidance, img_ids, txt_ids):
        B, S_img, _ = hidden.shape
        out = _empty_like_backend(hidden, (B, S_img, self.dit_cfg.out_channels))
        _ck(lib().flux2b_dit_forward_kv_cached(self._h, B, S_img, enc.shape[1], _ptr(hidden), _ptr(enc), _dtype_code(enc),
                                               _ptr(timestep), _ptr(guidance), _ptr(img_ids), _ptr(txt_ids), _ptr(out)))
        return out

    def block_output(self, index: int, S: int, D: int) -> np.ndarray:
        out = np.zeros((S, D), dtype=np.float32)
        _ck(lib().flux2b_get_block_output(self._h, index, _ptr(out), ctypes.c_size_t(out.nbytes)))
        return out

    # -- scheduler / latents
    def euler_step(self, sample, pred, sigma: float, sigma_next: float, pred_uncond=None, cfg: float = 1.0):
        n = int(np.prod(sample.shape))
        _ck(lib().flux2b_euler_step(self._h, _ptr(sample), _ptr(pred), _ptr(pred_uncond), ctypes.c_float(cfg),
                                    ctypes.c_float(sigma), ctypes.c_float(sigma_next), ctypes.c_size_t(n)))
        return sample

    def scale_noise(self, sample, noise, sigma: float):
        out = _empty_like_backend(sample, sample.shape)
        _ck(lib().flux2b_scale_noise(self._h, _ptr(sample), _ptr(noise), ctypes.c_float(sigma), _ptr(out),
                                     ctypes.c_size_t(int(np.prod(sample.shape)))))
        return out

    def repaint_blend(self, x, x0, eps, mask, sigma_next: float):
        _ck(lib().flux2b_repaint_blend(self._h, _ptr(x), _ptr(x0), _ptr(eps), _ptr(mask), ctypes.c_float(sigma_next),
                                       ctypes.c_size_t(int(np.prod(x.shape)))))
        return x

    def _perm(self, fn, x, out_shape, B, C, H, W):
        out = _empty_like_backend(x, out_shape)
        _ck(fn(self._h, _ptr(x), _ptr(out), B, C, H, W))
        return out

    def pack_patchified_to_sequence(self, x):
        B, C, H, W = x.shape
        return self._perm(lib().flux2b_pack_patchified_to_sequence, x, (B, H * W, C), B, C, H, W)

    def unpack_sequence_to_patchified(self, seq, height: int, width: int):
        B, _, C = seq.shape
        H, W = height // 16, width // 16
        return self._perm(lib().flux2b_unpack_sequence_to_patchified, seq, (B, C, H, W), B, C, H, W)

    def unpatchify_latents(self, x, latent_channels: int = 32):
        B, _, H, W = x.shape
        return self._perm(lib().flux2b_unpatchify_latents, x, (B, latent_channels, 2 * H, 2 * W), B, latent_channels, H, W)

    def pack_latents_to_patchified(self, x):
        B, C, H, W = x.shape
        return self._perm(lib().flux2b_pack_latents_to_patchified, x, (B, C * 4, H // 2, W // 2), B, C, H, W)

    def bn_latents(self, x, mean, var, eps: float = 1e-4, denormalize: bool = True):
        B, C, H, W = x.shape
        out = _empty_like_backend(x, x.shape)
        _ck(lib().flux2b_bn_latents(self._h, _ptr(x), _ptr(out), _ptr(mean), _ptr(var), ctypes.c_float(eps), B, C, H, W,
                                    int(denormalize)))
        return out

    # -- VAE
    def vae_decode(self, latents):
        B, _, h8, w8 = latents.shape
        out = _empty_like_backend(latents, (B, self.vae_cfg.out_channels, 8 * h8, 8 * w8))
        _ck(lib().flux2b_vae_decode(self._h, B, h8, w8, _ptr(latents), _ptr(out)))
        return out

    def vae_encode(self, image, noise=None):
        """AutoencoderKLFlux2.encode(_:samplePosterior:) (VAE/AutoencoderKL.swift:90-127). image [B,3,H,W] f32 in [-1,1];
        noise None = samplePosterior false, else the standard-normal noise to use -> latents [B, latent_ch, H/8, W/8]."""
        B, _, H, W = image.shape
        out = _empty_like_backend(image, (B, self.vae_cfg.latent_channels, H // 8, W // 8))
        _ck(lib().flux2b_vae_encode(self._h, B, H, W, _ptr(image), _ptr(noise), _ptr(out)))
        return out

    def encode_image_to_sequence(self, image, noise=None):
        """encodeImageToPackedSequence (Flux2Pipeline+ChainHelpers.swift:75-101) on a preprocessed image [B,3,H,W] f32:
        -> [B, (H/16)*(W/16), 4*latent_ch] BatchNorm-normalised packed latents."""
        B, _, H, W = image.shape
        out = _empty_like_backend(image, (B, (H // 16) * (W // 16), 4 * self.vae_cfg.latent_channels))
        _ck(lib().flux2b_encode_image_to_sequence(self._h, B, H, W, _ptr(image), _ptr(noise), _ptr(out)))
        return out

    def vae_decode_u8(self, latents) -> np.ndarray:
        B, _, h8, w8 = latents.shape
        out = np.zeros((B, 8 * h8, 8 * w8, 3), dtype=np.uint8)
        _ck(lib().flux2b_vae_decode_u8(self._h, B, h8, w8, _ptr(latents), _ptr(out)))
        return out

    # -- loop
    def _denoise_params(self, height, width, sigmas, enc, guidance, enc_uncond, cfg_scale, ref_latents, ref_ids, hook):
        sig = (ctypes.c_float * len(sigmas))(*sigmas)
        p = DenoiseParamsC()
        p.height, p.width, p.num_sigmas, p.sigmas = height, width, len(sigmas), sig
        g = np.array([guidance], dtype=np.float32) if guidance is not None else None
        p.guidance = _ptr(g)
        p.cfg_scale = cfg_scale
        p.enc, p.enc_uncond, p.enc_dtype, p.S_txt = _ptr(enc), _ptr(enc_uncond), _dtype_code(enc), enc.shape[-2]
        p.S_txt_uncond = int(enc_uncond.shape[-2]) if enc_uncond is not None else 0   # the negative prompt has its own length
        p.ref_latents, p.ref_ids = _ptr(ref_latents), _ptr(ref_ids)
        p.S_ref = int(ref_latents.shape[-2]) if ref_latents is not None else 0
        cb = None
        if hook is not None:
            def _cb(scp, lat, n, _user):
                sc = scp.contents
                arr = np.ctypeslib.as_array(lat, shape=(n,))
                try:
                    r = hook(sc, arr)
                    return 0 if (r is None or r == 0) else 1
                except Exception:
                    return 1
            cb = HOOK_T(_cb)
            p.hook = cb
        keep = (sig, g, cb, enc, enc_uncond, ref_latents, ref_ids)
        return p, keep

    def denoise(self, latents, enc, sigmas, height, width, guidance=None, enc_uncond=None, cfg_scale=1.0,
                ref_latents=None, ref_ids=None, hook: Optional[Callable] = None, kv_cache: bool = False):
        """In place on `latents` [1, S_img, 128]; hook(step_context, latents_view) may edit the view (Flux2StepHook).
        kv_cache=True with ref_latents: the klein-9b-kv loop (extract at step 0, cached afterwards)."""
        p, keep = self._denoise_params(height, width, sigmas, enc, guidance, enc_uncond, cfg_scale, ref_latents, ref_ids, hook)
        p.kv_cache = int(kv_cache)
        _ck(lib().flux2b_denoise(self._h, ctypes.byref(p), _ptr(latents)))
        return latents

    def generate(self, latents, enc, sigmas, height, width, rgb_out=None, **kw):
        p, keep = self._denoise_params(height, width, sigmas, enc, kw.get("guidance"), kw.get("enc_uncond"),
                                       kw.get("cfg_scale", 1.0), kw.get("ref_latents"), kw.get("ref_ids"), kw.get("hook"))
        if rgb_out is None:
            rgb_out = np.zeros((height, width, 3), dtype=np.uint8)
        _ck(lib().flux2b_generate(self._h, ctypes.byref(p), _ptr(latents), _ptr(rgb_out)))
        return rgb_out

    # -- profiler
    def prof_enable(self, on: bool = True):
        _ck(lib().flux2b_prof_enable(self._h, int(on)))

    def prof_reset(self):
        _ck(lib().flux2b_prof_reset(self._h))

    def prof_get(self, kind: int):
        ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        n = ctypes.c_int64()
        _ck(lib().flux2b_prof_get(self._h, kind, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl), ctypes.byref(by)))
        return {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}

    def launch_count(self) -> int:
        return int(lib().flux2b_launch_count(self._h))

    # -- single kernels
    def op_gemm(self, a16, w16, epilogue=0, bias=None, gate=None, res=None, cta_group=0, bn=0, out=None):
        M, K = a16.shape
        N = w16.shape[0]
        No = N // 2 if epilogue == 3 else N
        if out is None:
            import torch
            odt = torch.float32 if epilogue in (1, 2) else a16.dtype
            out = torch.empty((M, No), dtype=odt, device=a16.device)
        _ck(lib().flux2b_op_gemm(self._h, _ptr(a16), _ptr(w16), M, N, K, epilogue, _ptr(out), _ptr(bias), _ptr(gate),
                                 _ptr(res), cta_group, bn))
        return out

    def op_gemm_mx(self, quant, a16, w_packed, w_scales, return_quantized=False, bn=0, cta_group=0):
        """Native block-scaled GEMM. a16 [M,K] 16-bit torch tensor; w_packed uint32 [N, K*bits/32], w_scales uint8 [N, K/group]
        (the MLX layout of quant mode `quant` = "mxfp8" | "mxfp4" | "nvfp4") -> f32 [M, N]
        (+ quantised activation bytes [M, K*bits/8] and their scales [M, K/group] with return_quantized)."""
        import torch
        q = QUANT[quant] if isinstance(quant, str) else int(quant)
        bits, group = {3: (8, 32), 4: (4, 32), 5: (4, 16)}[q]
        M, K = a16.shape
        N = w_packed.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=a16.device)
        aq = np.zeros((M, K * bits // 8), dtype=np.uint8) if return_quantized else None
        sfa = np.zeros((M, K // group), dtype=np.uint8) if return_quantized else None
        _ck(lib().flux2b_op_gemm_mx(self._h, q, _ptr(a16), _ptr(w_packed), _ptr(w_scales), M, N, K, _ptr(out), _ptr(aq), _ptr(sfa), bn, cta_group))
        return (out, aq, sfa) if return_quantized else out

    def op_linear_quantized(self, quant, x16, w_packed, w_scales, w_biases=None, in_kernel=True, cta_group=0):
        """QuantizedLinear forward, W-only (x · dequant(W)^T). x16 [M, K] 16-bit torch tensor; w_packed uint32 [N, K*bits/32],
        w_scales / w_biases [N, K/group] as MLX holds them (f16 / bf16 for the affine modes, uint8 scales otherwise) -> f32 [M, N]."""
        import torch
        q = QUANT[quant] if isinstance(quant, str) else int(quant)
        M, K = x16.shape
        N = w_packed.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=x16.device)
        sbd = _dtype_code(w_scales) if w_biases is not None else F16
        _ck(lib().flux2b_op_linear_quantized(self._h, q, _ptr(x16), _ptr(w_packed), _ptr(w_scales), _ptr(w_biases), sbd, M, N, K,
                                             _ptr(out), int(in_kernel), cta_group))
        return out

    def op_gemm_mxfp8(self, a16, w_packed, w_scales, return_quantized=False):
        return self.op_gemm_mx(3, a16, w_packed, w_scales, return_quantized)

    def op_attention(self, qkv16, B, S, H, variant=0):
        import torch
        out = torch.empty((B * S, H * 128), dtype=qkv16.dtype, device=qkv16.device)
        _ck(lib().flux2b_op_attention(self._h, _ptr(qkv16), B, S, H, _ptr(out), variant))
        return out

    def op_attention_causal(self, qkv16, S, num_heads, num_kv_heads, key_lo=0, key_hi=0):
        import torch
        out = torch.empty((S, num_heads * 128), dtype=qkv16.dtype, device=qkv16.device)
        _ck(lib().flux2b_op_attention_causal(self._h, _ptr(qkv16), S, num_heads, num_kv_heads, key_lo, key_hi, _ptr(out)))
        return out

    def op_ln_modulate(self, x, shift, scale, out_dtype):
        import torch
        rows, D = x.shape
        out = torch.empty((rows, D), dtype=out_dtype, device=x.device)
        _ck(lib().flux2b_op_ln_modulate(self._h, _ptr(x), rows, D, _ptr(shift), _ptr(scale), _ptr(out)))
        return out

    def op_qk_norm_rope(self, qkv16, D, norm_q, norm_k, cos_t, sin_t):
        _ck(lib().flux2b_op_qk_norm_rope(self._h, _ptr(qkv16), qkv16.shape[0], D, _ptr(norm_q), _ptr(norm_k), _ptr(cos_t), _ptr(sin_t)))
        return qkv16

    def op_rope_table(self, ids):
        S = ids.shape[0]
        cos, sin = np.zeros((S, 128), np.float32), np.zeros((S, 128), np.float32)
        _ck(lib().flux2b_op_rope_table(self._h, _ptr(ids), S, _ptr(cos), _ptr(sin)))
        return cos, sin

    def op_timestep_embedding(self, t):
        out = np.zeros((t.shape[0], 256), np.float32)
        _ck(lib().flux2b_op_timestep_embedding(self._h, _ptr(t), t.shape[0], _ptr(out)))
        return out

    def op_conv2d(self, x16, w16, bias, res16=None, cta_group=0):
        import torch
        B, H, W, Cin = x16.shape
        Cout, k = w16.shape[0], w16.shape[1]
        out = torch.empty((B, H, W, Cout), dtype=x16.dtype, device=x16.device)
        _ck(lib().flux2b_op_conv2d(self._h, _ptr(x16), _ptr(w16), _ptr(bias), _ptr(res16), _ptr(out), B, H, W, Cin, Cout, k, cta_group))
        return out

    def op_groupnorm_silu(self, x16, gamma, beta, G, eps, silu=True):
        import torch
        B, H, W, C = x16.shape
        out = torch.empty_like(x16)
        _ck(lib().flux2b_op_groupnorm_silu(self._h, _ptr(x16), _ptr(out), _ptr(gamma), _ptr(beta), B, H * W, C, G,
                                           ctypes.c_float(eps), int(silu)))
        return out


class TextEncoder(Context):
    """A text-encoder context (flux2b_te_create): Qwen3Model / MistralModel as the embedding extractors use them
    (FluxTextEncoders/Model/Qwen3/Qwen3Model.swift:104-191). Weights go in under the HF / Swift module paths with the
    inherited set_tensor / load_weights / load_safetensors / finalize."""

    def __init__(self, cfg, quant: int = 0, device: int = 0, options: Optional[Dict[str, int]] = None):
        L = lib()
        self._h = ctypes.c_void_p()
        self.te_cfg = cfg
        self.dit_cfg = self.vae_cfg = None
        tc = TeConfigC(cfg.vocab_size, cfg.hidden_size, cfg.intermediate_size, cfg.num_layers, cfg.num_heads, cfg.num_kv_heads,
                       cfg.head_dim, int(cfg.qk_norm), cfg.rms_norm_eps, cfg.rope_theta, int(getattr(cfg, "max_position_embeddings", 0)))
        _ck(L.flux2b_te_create(device, ctypes.byref(tc), quant, ctypes.byref(self._h)))
        for k, v in (options or {}).items():
            self.set_option(k, v)
        self._keep = []

    def forward_with_hidden_states(self, input_ids, layer_indices: Sequence[int], attention_mask=None, out_dtype=F32):
        """Qwen3Model.forwardWithHiddenStates(_:layerIndices:attentionMask:) followed by the concatenation along the hidden
        axis (KleinEmbeddingExtractor.swift:98-121). input_ids / attention_mask: int32 [B, S] -> [B, S, n * hidden]."""
        ids = np.ascontiguousarray(np.asarray(input_ids, dtype=np.int32))
        assert ids.ndim == 2
        B, S = ids.shape
        m = None if attention_mask is None else np.ascontiguousarray(np.asarray(attention_mask, dtype=np.int32).reshape(B, S))
        li = (ctypes.c_int * len(layer_indices))(*[int(i) for i in layer_indices])
        np_dt = {F32: np.float32, F16: np.float16, BF16: np.uint16}[out_dtype]   # bf16 comes back as raw uint16 words
        out = np.empty((B, S, len(layer_indices) * self.te_cfg.hidden_size), dtype=np_dt)
        _ck(lib().flux2b_te_hidden_states(self._h, B, S, _ptr(ids), _ptr(m), li, len(layer_indices), _ptr(out), out_dtype))
        return out


class KleinEmbeddingExtractor:
    """Device half of KleinEmbeddingExtractor.extractKleinEmbeddings (Embeddings/KleinEmbeddingExtractor.swift:46-133): takes
    the token ids the Swift side produced (chat template + tokenizer stay on the host), truncates, RIGHT-pads with
    <|endoftext|> (151643) to 512, builds the mask and returns the [1, 512, 3 * hidden] embeddings of layers 9 / 18 / 27."""
    PAD_TOKEN_ID = 151643
    HIDDEN_STATE_LAYERS = (9, 18, 27)     # Embeddings/KleinConfig.swift:28-31
    MAX_SEQUENCE_LENGTH = 512             # Embeddings/KleinConfig.swift:120

    def __init__(self, model: TextEncoder):
        self.model = model

    def extract(self, token_ids: Sequence[int], max_length: int = MAX_SEQUENCE_LENGTH, out_dtype=F32):
        ids = list(token_ids)[:max_length]
        n = len(ids)
        ids = ids + [self.PAD_TOKEN_ID] * (max_length - n)
        mask = [1] * n + [0] * (max_length - n)
        return self.model.forward_with_hidden_states([ids], self.HIDDEN_STATE_LAYERS, [mask], out_dtype)


class FluxEmbeddingExtractor:
    """Device half of EmbeddingExtractor.extractFluxEmbeddings (Embeddings/EmbeddingExtractor.swift:202-285; Flux.2 Dev, Mistral Small
    3.2): token ids of the chat-templated prompt (system message + user prompt, TekkenTokenizer on the Swift side) are truncated,
    LEFT-padded with the tokenizer's pad token to 512, masked, and the hidden states of layers 10 / 20 / 30 (hidden_states index i =
    output of decoder layer i, 0 = embeddings) are concatenated to [1, 512, 3 * 5120 = 15360]."""
    HIDDEN_STATE_LAYERS = (10, 20, 30)    # FluxConfig.hiddenStateLayers
    MAX_SEQUENCE_LENGTH = 512             # FluxConfig.maxSequenceLength

    def __init__(self, model: "TextEncoder", pad_token_id: int):
        self.model = model
        self.pad_token_id = int(pad_token_id)

    def extract(self, token_ids: Sequence[int], max_length: int = MAX_SEQUENCE_LENGTH, out_dtype=F32):
        ids = list(token_ids)[:max_length]
        n = len(ids)
        ids = [self.pad_token_id] * (max_length - n) + ids
        mask = [0] * (max_length - n) + [1] * n
        return self.model.forward_with_hidden_states([ids], self.HIDDEN_STATE_LAYERS, [mask], out_dtype)


def _empty_like_backend(x, shape):
    if _is_torch(x):
        import torch
        return torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=x.device)
    return np.zeros(tuple(int(s) for s in shape), dtype=np.float32)


# ------------------------------------------------------------------------------------------------ reference-shaped mirrors
class Flux2Transformer2DModel:
    """Mirror of Flux2Transformer2DModel (Flux2Transformer.swift:22): callAsFunction -> the C ABI."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def __call__(self, hidden_states, encoder_hidden_states, timestep, guidance=None, img_ids=None, txt_ids=None):
        return self.ctx.dit_forward(hidden_states, encoder_hidden_states, timestep, guidance, img_ids, txt_ids)

    def forward_kv_extract(self, hidden_states, reference_hidden_states, encoder_hidden_states, timestep, guidance,
                           img_ids, ref_ids, txt_ids):
        return self.ctx.dit_forward_kv_extract(hidden_states, reference_hidden_states, encoder_hidden_states, timestep,
                                               guidance, img_ids, ref_ids, txt_ids)

    def forward_kv_cached(self, hidden_states, encoder_hidden_states, timestep, guidance, img_ids, txt_ids):
        return self.ctx.dit_forward_kv_cached(hidden_states, encoder_hidden_states, timestep, guidance, img_ids, txt_ids)


class AutoencoderKLFlux2:
    """Mirror of AutoencoderKLFlux2 (VAE/AutoencoderKL.swift:46): decode only (encoder is SURVEY §8f 'next')."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def decode(self, z):
        return self.ctx.vae_decode(z)
