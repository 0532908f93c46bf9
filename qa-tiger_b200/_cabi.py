"""ctypes binding of include/qat.h.  There is NO fallback: if libqat_b200.so is missing or an entry
point returns non-zero, a RuntimeError is raised ("errors, never fallbacks")."""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from typing import Dict, Optional, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QAT_LIB") or os.path.join(HERE, "csrc", "libqat_b200.so")   # QAT_LIB: A/B builds while profiling

ABI_VERSION = 4
PREC_FP32, PREC_BF16 = 0, 1
DT_F32, DT_BF16 = 0, 1
OP_LINEAR, OP_AVQ, OP_PATCH, OP_MOE, OP_GROUND, OP_TOME_MATCH, OP_TOME_MERGE, OP_TOME_CHAIN, OP_CLIP_TEXT, OP_FORWARD, OP_TOME_BLOCK, OP_TOME_ATTENTION = range(12)
PREC_NAMES = {"fp32": PREC_FP32, "bf16": PREC_BF16}

c_f32p = C.c_void_p  # all device pointers travel as void*


class LinearW(C.Structure):
    _fields_ = [("w", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_f16", C.c_void_p), ("bias", C.c_void_p),
                ("n", C.c_int32), ("k", C.c_int32)]


class LayerNormW(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p)]


class LinearArgs(C.Structure):
    _fields_ = [("rows", C.c_int32), ("prec", C.c_int32), ("stream", C.c_void_p), ("x", C.c_void_p),
                ("w", LinearW), ("y", C.c_void_p), ("relu", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t)]


class GemmArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("M", C.c_int32), ("stream", C.c_void_p), ("a", C.c_void_p), ("a_hi", C.c_void_p),
                ("a_lo", C.c_void_p), ("a_f16", C.c_void_p), ("lda", C.c_int32), ("w", LinearW), ("relu", C.c_int32),
                ("res", C.c_void_p), ("res_f16", C.c_void_p), ("ldr", C.c_int32), ("c", C.c_void_p), ("c_hi", C.c_void_p),
                ("c_lo", C.c_void_p), ("c_f16", C.c_void_p), ("ldc", C.c_int32),
                ("coef", C.c_void_p), ("coef_E", C.c_int32), ("coef_T", C.c_int32), ("coef_Ch", C.c_int32), ("coef_mod", C.c_int64)]


class AvqArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("L", C.c_int32), ("D", C.c_int32), ("H", C.c_int32),
                ("prec", C.c_int32), ("stream", C.c_void_p),
                ("audio", C.c_void_p), ("video", C.c_void_p), ("words", C.c_void_p),
                ("in_src", LinearW), ("in_words", LinearW), ("out_cat", LinearW), ("ffn1", LinearW), ("ffn2", LinearW),
                ("norm1", LayerNormW), ("norm2", LayerNormW),
                ("out_audio", C.c_void_p), ("out_video", C.c_void_p), ("attw_audio", C.c_void_p), ("attw_video", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class PatchArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("P", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("Cin", C.c_int32),
                ("prec", C.c_int32), ("stream", C.c_void_p),
                ("patch", C.c_void_p), ("patch_f16", C.c_void_p), ("audio", C.c_void_p), ("video", C.c_void_p),
                ("has_proj", C.c_int32),
                ("proj", LinearW), ("slf_in", LinearW), ("slf_out", LinearW), ("crs_q", LinearW), ("crs_kv", LinearW),
                ("crs_out", LinearW), ("mlp0", LinearW), ("mlp2", LinearW),
                ("anorm", LayerNormW), ("vnorm", LayerNormW),
                ("out_a", C.c_void_p), ("out_v", C.c_void_p), ("chunk_frames", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class GaussArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("E", C.c_int32), ("K", C.c_int32), ("C", C.c_int32),
                ("dtype", C.c_int32), ("sigma", C.c_float), ("stream", C.c_void_p),
                ("experts", C.c_void_p), ("stride_t", C.c_int64), ("stride_b", C.c_int64),
                ("router_logits", C.c_void_p), ("gauss_raw", C.c_void_p), ("out", C.c_void_p),
                ("topk_inds", C.c_void_p), ("topk_probs", C.c_void_p), ("gauss_weight", C.c_void_p),
                ("t_split", C.c_int32)]


class MoeArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("E", C.c_int32), ("K", C.c_int32),
                ("prec", C.c_int32), ("sigma", C.c_float), ("stream", C.c_void_p),
                ("qst", C.c_void_p), ("data", C.c_void_p), ("sub_a", C.c_void_p), ("sub_v", C.c_void_p),
                ("attn_q", LinearW), ("attn_kv", LinearW), ("attn_rg", LinearW),
                ("w1_cat", LinearW), ("w2_ext", LinearW), ("norm_a", LayerNormW), ("norm_v", LayerNormW),
                ("out_a", C.c_void_p), ("out_v", C.c_void_p),
                ("topk_inds", C.c_void_p), ("topk_probs", C.c_void_p), ("gauss_weight", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class GroundArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("prec", C.c_int32), ("stream", C.c_void_p),
                ("qst", C.c_void_p), ("d0", C.c_void_p), ("d1", C.c_void_p), ("data", C.c_void_p), ("S", C.c_int32),
                ("attn_q", LinearW), ("attn_kv", LinearW), ("attn_out", LinearW), ("mlp0", LinearW), ("mlp2", LinearW),
                ("norm", LayerNormW), ("out", C.c_void_p), ("has_head", C.c_int32), ("head", LinearW),
                ("logits", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class ForwardArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("L", C.c_int32), ("P", C.c_int32), ("prec", C.c_int32), ("stream", C.c_void_p),
                ("quest", C.c_void_p), ("words", C.c_void_p), ("audio", C.c_void_p), ("video", C.c_void_p), ("patch", C.c_void_p),
                ("audio_proj", LinearW), ("video_proj", LinearW), ("words_proj", LinearW), ("quest_proj", LinearW),
                ("avq", AvqArgs), ("patch_sel", PatchArgs), ("at", MoeArgs), ("vt", MoeArgs), ("ground", GroundArgs),
                ("q_cat", LinearW), ("logits", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class TomeMatchArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("t", C.c_int32), ("C", C.c_int32), ("r", C.c_int32), ("class_token", C.c_int32),
                ("stream", C.c_void_p), ("metric", C.c_void_p), ("unm_idx", C.c_void_p), ("src_idx", C.c_void_p),
                ("dst_idx", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class TomeMergeArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("t", C.c_int32), ("C", C.c_int32), ("r", C.c_int32), ("mode", C.c_int32), ("distill_token", C.c_int32), ("stream", C.c_void_p),
                ("x", C.c_void_p), ("size", C.c_void_p), ("unm_idx", C.c_void_p), ("src_idx", C.c_void_p),
                ("dst_idx", C.c_void_p), ("x_out", C.c_void_p), ("size_out", C.c_void_p), ("group", C.c_void_p),
                ("T0", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class TomeChainArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("t0", C.c_int32), ("C", C.c_int32), ("n_steps", C.c_int32),
                ("r_list", C.POINTER(C.c_int32)), ("class_token", C.c_int32), ("stream", C.c_void_p),
                ("x", C.c_void_p), ("x_out", C.c_void_p), ("size_out", C.c_void_p), ("group", C.c_void_p),
                ("idx_trace", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class TomeBlockArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("t", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("r", C.c_int32), ("class_token", C.c_int32),
                ("prop_attn", C.c_int32), ("prec", C.c_int32), ("eps", C.c_float), ("stream", C.c_void_p),
                ("x", C.c_void_p), ("size", C.c_void_p), ("norm1", LayerNormW), ("norm2", LayerNormW),
                ("qkv", LinearW), ("proj", LinearW), ("fc1", LinearW), ("fc2", LinearW),
                ("x_out", C.c_void_p), ("size_out", C.c_void_p), ("metric", C.c_void_p),
                ("unm_idx", C.c_void_p), ("src_idx", C.c_void_p), ("dst_idx", C.c_void_p), ("group", C.c_void_p), ("T0", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class AttnArgs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nseq", C.c_int32), ("H", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32),
                ("scale", C.c_float), ("causal", C.c_int32), ("stream", C.c_void_p),
                ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
                ("q_sn", C.c_int64), ("q_si", C.c_int64), ("k_sn", C.c_int64), ("k_si", C.c_int64),
                ("k_shift", C.c_int32), ("k_mod", C.c_int32), ("o", C.c_void_p), ("o_f16", C.c_void_p),
                ("o_rn", C.c_int64), ("o_ri", C.c_int64), ("ldo", C.c_int32), ("attw", C.c_void_p), ("kbias", C.c_void_p)]


class GemmAttnArgs(C.Structure):
    _fields_ = [("cross", C.c_int32), ("nframes", C.c_int32), ("P", C.c_int32), ("H", C.c_int32), ("K", C.c_int32),
                ("scale", C.c_float), ("stream", C.c_void_p), ("x_f16", C.c_void_p), ("ldx", C.c_int32),
                ("w_f16", C.c_void_p), ("bias", C.c_void_p), ("q_f16", C.c_void_p), ("o_f16", C.c_void_p), ("ldo", C.c_int32)]


class ClipLayerW(C.Structure):
    _fields_ = [("ln_1", LayerNormW), ("ln_2", LayerNormW), ("in_proj", LinearW), ("out_proj", LinearW),
                ("c_fc", LinearW), ("c_proj", LinearW)]


class ClipTextArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("E", C.c_int32),
                ("n_layers", C.c_int32), ("vocab", C.c_int32), ("prec", C.c_int32), ("stream", C.c_void_p),
                ("tokens", C.c_void_p), ("token_embedding", C.c_void_p), ("positional_embedding", C.c_void_p),
                ("layers", C.POINTER(ClipLayerW)), ("ln_final", LayerNormW), ("text_projection", LinearW),
                ("quest", C.c_void_p), ("words", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


GATHER_MAX = 4


class GatherArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_tensors", C.c_int32), ("n_slots", C.c_int32), ("stream", C.c_void_p), ("slots", C.c_void_p),
                ("src", C.c_void_p * GATHER_MAX), ("dst", C.c_void_p * GATHER_MAX), ("row_bytes", C.c_size_t * GATHER_MAX),
                ("src_stride", C.c_size_t * GATHER_MAX)]


EXPORTS = {
    "qat_gemm": (C.c_int, [C.POINTER(GemmArgs)]),
    "qat_attention_fwd": (C.c_int, [C.POINTER(AttnArgs)]),
    "qat_gemm_attn_fwd": (C.c_int, [C.POINTER(GemmAttnArgs)]),
    "qat_linear_fwd": (C.c_int, [C.POINTER(LinearArgs)]),
    "qat_avq_cross_attn_fwd": (C.c_int, [C.POINTER(AvqArgs)]),
    "qat_patch_select_fwd": (C.c_int, [C.POINTER(PatchArgs)]),
    "qat_gauss_reduce_fwd": (C.c_int, [C.POINTER(GaussArgs)]),
    "qat_temp_moe_fwd": (C.c_int, [C.POINTER(MoeArgs)]),
    "qat_qst_grounding_head_fwd": (C.c_int, [C.POINTER(GroundArgs)]),
    "qat_clip_text_fwd": (C.c_int, [C.POINTER(ClipTextArgs)]),
    "qat_forward_fwd": (C.c_int, [C.POINTER(ForwardArgs)]),
    "qat_tome_match": (C.c_int, [C.POINTER(TomeMatchArgs)]),
    "qat_tome_clamp_r": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "qat_tome_merge": (C.c_int, [C.POINTER(TomeMergeArgs)]),
    "qat_tome_chain": (C.c_int, [C.POINTER(TomeChainArgs)]),
    "qat_tome_chain_t_end": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "qat_tome_chain_trace_len": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "qat_workspace_bytes": (C.c_size_t, [C.c_int, C.c_void_p]),
    "qat_last_error": (C.c_char_p, []),
    "qat_abi_version": (C.c_int, []),
    "qat_launch_count": (C.c_longlong, [C.c_int]),
    "qat_gather_rows": (C.c_int, [C.POINTER(GatherArgs)]),
    "qat_tome_block_fwd": (C.c_int, [C.POINTER(TomeBlockArgs)]),
    "qat_tome_attention_fwd": (C.c_int, [C.POINTER(TomeBlockArgs)]),
    "qat_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "qat_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load libqat_b200.so (once).  Missing library => RuntimeError, never a silent fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                        "(qa-tiger_b200 has no CPU or PyTorch fallback for its CUDA path)")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in EXPORTS.items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                if l.qat_abi_version() != ABI_VERSION:
                    raise RuntimeError("libqat_b200.so ABI version mismatch")
                _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().qat_last_error().decode()}")


def set_option(name: str, value: int) -> None:
    """A/B switch of the library (include/qat.h `qat_set_option`): process-wide, takes effect for calls entered afterwards."""
    check(lib().qat_set_option(name.encode(), int(value)), "qat_set_option")


def get_option(name: str) -> int:
    v = C.c_int(0)
    check(lib().qat_get_option(name.encode(), C.byref(v)), "qat_get_option")
    return v.value


@contextlib.contextmanager
def option(name: str, value: int):
    """with option("gemm_cg", 1): ...  -- set an A/B switch for the calls inside the block, restore it afterwards."""
    prev = get_option(name)
    set_option(name, value)
    try:
        yield
    finally:
        set_option(name, prev)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda_f32(t: torch.Tensor, name: str, shape: Optional[Tuple[int, ...]] = None, allow_f16: bool = False):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (qa-tiger_b200 has no CPU path), got {t.device}")
    if t.dtype != torch.float32 and not (allow_f16 and t.dtype == torch.float16):
        raise RuntimeError(f"{name}: expected float32{' or float16' if allow_f16 else ''}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor (SURVEY 8b: the shim asserts, it does not copy)")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ---- weights ------------------------------------------------------------------------------------
def aligned(t: torch.Tensor) -> torch.Tensor:
    """A detached contiguous fp32 tensor on a 16-byte boundary (the kernels read weights with 128-bit loads / TMA).  Module
    parameters normally are; nn.DataParallel's replicas are views into coalesced broadcast buffers and need not be -- those are
    copied (the copy lives in the prepared-weight cache, keyed on the replica's data_ptr / version)."""
    t = t.detach().contiguous().float()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class Lin:
    """A prepared nn.Linear operand: fp32 weight (+ bf16 hi/lo split when prec == bf16) and bias.
    Holds the tensors alive; `.c` is the C struct."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], prec: int, f16: bool = False):
        w = aligned(w)
        self.w = w
        self.b = None if b is None else aligned(b)
        self.hi = self.lo = self.h = None
        if prec == PREC_BF16:
            self.hi = w.to(torch.bfloat16)
            self.lo = (w - self.hi.float()).to(torch.bfloat16)
            if f16:
                self.h = w.to(torch.float16)
        self.c = LinearW(ptr(w), ptr(self.hi), ptr(self.lo), ptr(self.h), ptr(self.b), w.shape[0], w.shape[1])


class LN:
    def __init__(self, m: torch.nn.LayerNorm):
        self.g = aligned(m.weight)
        self.b = aligned(m.bias)
        self.c = LayerNormW(ptr(self.g), ptr(self.b))


# Prepared (derived) weights -- concatenated in-projections, bf16 hi/lo and fp16 planes, W2ext -- are cached per module,
# keyed on (epoch, data_ptr, _version) of every parameter.  `_version` moves on in-place tensor ops (`p.mul_()`, optimizer
# steps) and `data_ptr` on `.to()` / load_state_dict into fresh storage, but writes through `.data` (`p.data.copy_(...)`,
# EMA weight swaps) change NEITHER: call `invalidate_prepared()` after such writes.  nn.Module.load_state_dict and
# `.to()/.cuda()/.half()` of the mirrored modules call it automatically (`_QatModule._apply` / `_load_from_state_dict`).
_prepared_epoch = 0


def invalidate_prepared() -> None:
    """Drop every cached derived weight of every module (cheap: they are rebuilt lazily on the next forward)."""
    global _prepared_epoch
    _prepared_epoch += 1


def params_version(module: torch.nn.Module):
    return (_prepared_epoch,) + tuple((p.data_ptr(), p._version) for p in module.parameters(recurse=True))


GEMM_SIMT, GEMM_TF32, GEMM_BF16X3, GEMM_F16 = 0, 1, 2, 3


def split_planes(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def gemm(kind: int, a: torch.Tensor, lin: "Lin", relu: bool = False, res: Optional[torch.Tensor] = None,
         out_planes: bool = False, a_planes=None, out_f16: bool = False):
    """One contraction C = act(A W^T + b) + res on the chosen engine (parity tests / roofline timing).
    out_f16: the only output is an fp16 plane (TMA-store epilogue); `res` may then be an fp16 tensor."""
    M, K = a.shape
    N = lin.w.shape[0]
    dev = a.device
    g = GemmArgs()
    g.kind, g.M, g.stream, g.lda, g.w, g.relu, g.ldc = kind, M, stream_ptr(dev), a.stride(0), lin.c, int(relu), N
    hold = []
    if kind == GEMM_BF16X3:
        hi, lo = a_planes if a_planes is not None else split_planes(a)
        hold += [hi, lo]
        g.a_hi, g.a_lo = hi.data_ptr(), lo.data_ptr()
    elif kind == GEMM_F16:
        ah = a if a.dtype == torch.float16 else a.to(torch.float16)
        hold.append(ah)
        g.a_f16, g.lda = ah.data_ptr(), ah.stride(0)
    else:
        g.a = a.data_ptr()
    if res is not None:
        if res.dtype == torch.float16:
            g.res_f16, g.ldr = res.data_ptr(), res.stride(0)
        else:
            g.res, g.ldr = res.data_ptr(), res.stride(0)
    if out_f16:
        c = torch.empty(M, N, device=dev, dtype=torch.float16)
        g.c_f16 = c.data_ptr()
        with torch.cuda.device(dev):
            check(lib().qat_gemm(C.byref(g)), "qat_gemm")
        return c
    c = torch.empty(M, N, device=dev, dtype=torch.float32)
    g.c = c.data_ptr()
    planes = None
    if out_planes:
        planes = (torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16))
        g.c_hi, g.c_lo = planes[0].data_ptr(), planes[1].data_ptr()
    with torch.cuda.device(dev):
        check(lib().qat_gemm(C.byref(g)), "qat_gemm")
    return (c, planes) if out_planes else c


def gemm_reduce(a: torch.Tensor, lin: "Lin", coef: torch.Tensor, T: int, Ch: int, relu: bool = True, a_planes=None) -> torch.Tensor:
    """The expert first-layer GEMM with M4 in its epilogue (bf16x3 kind): a [S*T, K] rows = (sample, t), coef [R, E] with row
    (s*T + t) % R -> out [S, N] = sum_t coef[., n // Ch] * act(a W^T + b)   (parity tests / roofline timing)."""
    M, K = a.shape
    N = lin.w.shape[0]
    dev = a.device
    g = GemmArgs()
    g.kind, g.M, g.stream, g.lda, g.w, g.relu, g.ldc = GEMM_BF16X3, M, stream_ptr(dev), a.stride(0), lin.c, int(relu), N
    hi, lo = a_planes if a_planes is not None else split_planes(a)
    g.a_hi, g.a_lo = hi.data_ptr(), lo.data_ptr()
    coef = coef.contiguous().float()
    g.coef, g.coef_E, g.coef_T, g.coef_Ch, g.coef_mod = coef.data_ptr(), coef.shape[1], T, Ch, coef.shape[0]
    c = torch.empty(M // T, N, device=dev, dtype=torch.float32)
    g.c = c.data_ptr()
    with torch.cuda.device(dev):
        check(lib().qat_gemm(C.byref(g)), "qat_gemm")
    return c


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False, need_weights: bool = False,
              key_bias: Optional[torch.Tensor] = None):
    """softmax(q k^T / 8) v on the library's attention cores (parity tests / roofline timing).
    fp32: q [N,Lq,H,64], k, v [N,Lk,H,64] (contiguous) -> o [N,Lq,H*64] (+ head-averaged weights [N,Lq,Lk]).
    fp16: q [N,Lq,H,64], k and v are the halves of ONE packed tensor kv [N,Lk,2,H,64] (pass kv as `k`, v=None) -> fp16 o."""
    N, Lq, H, hd = q.shape
    assert hd == 64
    a = AttnArgs()
    a.nseq, a.H, a.Lq, a.scale, a.causal, a.stream = N, H, Lq, 0.125, int(causal), stream_ptr(q.device)
    if q.dtype == torch.float16:
        Lk = k.shape[1]
        o = torch.empty(N, Lq, H * 64, device=q.device, dtype=torch.float16)
        a.kind, a.Lk, a.q, a.k, a.q_si, a.k_si, a.o_f16, a.ldo = 1, Lk, q.data_ptr(), k.data_ptr(), H * 64, 2 * H * 64, o.data_ptr(), H * 64
        with torch.cuda.device(q.device):
            check(lib().qat_attention_fwd(C.byref(a)), "qat_attention_fwd")
        return o
    Lk = k.shape[1]
    o = torch.empty(N, Lq, H * 64, device=q.device, dtype=torch.float32)
    w = torch.zeros(N, Lq, Lk, device=q.device, dtype=torch.float32) if need_weights else None
    a.kind, a.Lk, a.q, a.k, a.v = 0, Lk, q.data_ptr(), k.data_ptr(), v.data_ptr()
    a.q_sn, a.q_si, a.k_sn, a.k_si, a.k_shift, a.k_mod = Lq * H * 64, H * 64, Lk * H * 64, H * 64, 0, N
    a.o, a.o_rn, a.o_ri, a.ldo, a.attw = o.data_ptr(), Lq, 1, H * 64, ptr(w)
    if key_bias is not None:
        require_cuda_f32(key_bias, "key_bias", (N, Lk))
        a.kbias = key_bias.data_ptr()
    with torch.cuda.device(q.device):
        check(lib().qat_attention_fwd(C.byref(a)), "qat_attention_fwd")
    return (o, w) if need_weights else o


def gemm_attn(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, P: int, H: int, q: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PatchSelecter's fused projection + attention kernels (parity tests / roofline timing): x [nframes*P, K] fp16, w fp16
    in_proj ([3D, K]; with `q` [nframes, 2, D] fp16: its k | v rows [2D, K]), bias fp32 -> o fp16 [nframes*P, D] ([nframes*2, D])."""
    rows, K = x.shape
    nf, D = rows // P, H * 64
    assert rows == nf * P and x.dtype == torch.float16 and w.dtype == torch.float16 and bias.dtype == torch.float32
    assert x.is_contiguous() and w.is_contiguous() and bias.is_contiguous() and (q is None or (q.is_contiguous() and q.dtype == torch.float16))
    o = torch.empty(nf * (P if q is None else 2), D, device=x.device, dtype=torch.float16)
    a = GemmAttnArgs()
    a.cross, a.nframes, a.P, a.H, a.K, a.scale, a.stream = int(q is not None), nf, P, H, K, 0.125, stream_ptr(x.device)
    a.x_f16, a.ldx, a.w_f16, a.bias, a.q_f16, a.o_f16, a.ldo = x.data_ptr(), K, w.data_ptr(), bias.data_ptr(), ptr(q), o.data_ptr(), D
    with torch.cuda.device(x.device):
        check(lib().qat_gemm_attn_fwd(C.byref(a)), "qat_gemm_attn_fwd")
    return o


# ---- workspace ----------------------------------------------------------------------------------
_ws: Dict[Tuple[int, int, int], torch.Tensor] = {}
_ws_retired: list = []          # superseded buffers that a captured CUDA graph may still point into: never released
_ws_lock = threading.Lock()     # nn.DataParallel drives one Python thread per GPU through this module


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per (device, stream, host thread) grow-only scratch buffer owned by torch's caching allocator.  A buffer that is outgrown
    while a stream capture is (or may have been) using it is retired, not freed: a captured graph has its address baked in.
    The thread is part of the key because two host threads may legally issue on ONE stream (nn.DataParallel with replicas on
    the same device, the legacy default stream): their calls interleave in stream order and must not share scratch memory."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device), threading.get_ident())
    with _ws_lock:
        buf = _ws.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and (getattr(buf, "_qat_captured", False) or torch.cuda.is_current_stream_capturing()):
                _ws_retired.append(buf)
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            _ws[key] = buf
        if torch.cuda.is_current_stream_capturing():
            buf._qat_captured = True
        return buf


def run(op: int, fn_name: str, args, device: torch.device):
    """Size the workspace from the same args struct, attach it, call the entry point."""
    l = lib()
    need = l.qat_workspace_bytes(op, C.addressof(args))
    if need == 0:
        raise RuntimeError(f"{fn_name}: workspace sizing failed: {l.qat_last_error().decode()}")
    ws = workspace(device, need)
    args.workspace = ws.data_ptr()
    args.workspace_bytes = ws.numel()
    args.stream = stream_ptr(device)
    with torch.cuda.device(device):
        check(getattr(l, fn_name)(C.byref(args)), fn_name)
