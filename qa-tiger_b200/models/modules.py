"""Drop-in mirrors of the reference's src/models/modules.py blocks, backed by libqat_b200 (sm_100a).

Same class names, constructor arguments, sub-module / parameter names (so `best.pt` loads with the
reference's own keys), construction order (same RNG stream => same random init) and forward
signatures / return shapes.  The forward bodies contain no math: they pack device pointers and call
one C-ABI entry point each (include/qat.h).  Eval only, CUDA only, no fallback.

Reference: /root/reference/src/models/modules.py
  Projection :19-43 | AVQCrossAttn :159-265 | QstGrounding :268-347 | TempMoE :350-612 | PatchSelecter :615-716
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from .. import _cabi as cabi

_DEFAULT_PREC = "bf16"


def set_default_precision(name: str):
    """'fp32' (CUDA-core FFMA, 1e-5 gate) or 'bf16' (tcgen05 tensor cores, 1e-3 gate)."""
    global _DEFAULT_PREC
    if name not in cabi.PREC_NAMES:
        raise ValueError(f"precision must be one of {list(cabi.PREC_NAMES)}")
    _DEFAULT_PREC = name


def get_default_precision() -> str:
    return _DEFAULT_PREC


class _QatModule(nn.Module):
    """Shared plumbing: precision selection, eval-only guard, cached weight preparation."""

    precision: Optional[str] = None  # None -> module default -> global default

    def _prec(self) -> int:
        return cabi.PREC_NAMES[self.precision or _DEFAULT_PREC]

    def _guard(self):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: the B200 path is inference-only; call .eval() "
                               "(the reference applies dropout in train mode, this path has none)")

    # derived weights are dropped whenever the parameters may have changed behind the (data_ptr, _version) key:
    # .to() / .cuda() / .float() (-> _apply) and load_state_dict (-> _load_from_state_dict); `.data` writes need an
    # explicit qa_tiger_b200.invalidate_prepared() (see _cabi.params_version)
    def _apply(self, fn, *args, **kwargs):
        cabi.invalidate_prepared()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        cabi.invalidate_prepared()
        return super()._load_from_state_dict(*args, **kwargs)

    def _prepared(self, prec: int):
        key = (prec, cabi.params_version(self))
        cache = self.__dict__.get("_qat_cache")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                cache = (key, self._prepare(prec))
            self.__dict__["_qat_cache"] = cache
        return cache[1]

    def _prepare(self, prec: int):  # pragma: no cover - overridden
        raise NotImplementedError


def _init_linear(m: nn.Module):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


# ------------------------------------------------------------------------------------------------
class Projection(_QatModule):
    """modules.py:19-43: y = proj(x) over the last dim."""

    def __init__(self, inp_dim: int = 512, d_model: int = 512):
        super().__init__()
        self.proj = nn.Linear(inp_dim, d_model)

    def _prepare(self, prec):
        return cabi.Lin(self.proj.weight, self.proj.bias, prec, f16=True)   # fp16 copy: patch_proj on an fp16 feature store

    def forward(self, inp: Tensor) -> Tensor:
        self._guard()
        prec = self._prec()
        w = self._prepared(prec)
        cabi.require_cuda_f32(inp, "Projection input")
        rows = inp.numel() // inp.shape[-1]
        out = torch.empty(*inp.shape[:-1], w.w.shape[0], device=inp.device, dtype=torch.float32)
        a = cabi.LinearArgs()
        a.rows, a.prec, a.x, a.w, a.y, a.relu = rows, prec, inp.data_ptr(), w.c, out.data_ptr(), 0
        cabi.run(cabi.OP_LINEAR, "qat_linear_fwd", a, inp.device)
        return out


# ------------------------------------------------------------------------------------------------
class AVQCrossAttn(_QatModule):
    """modules.py:159-265: question / self / cross attention on audio and video, both directions."""

    def __init__(self, d_model: int = 512, nhead: int = 8, dropout: float = 0.1):
        super().__init__()
        self.qst_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)   # parameter holders only
        self.crs_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.slf_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, d_model)
        self.linear2 = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        nn.init.kaiming_normal_(self.linear1.weight)
        nn.init.constant_(self.linear1.bias, 0)
        nn.init.kaiming_normal_(self.linear2.weight)
        nn.init.constant_(self.linear2.bias, 0)
        self.d_model, self.nhead = d_model, nhead

    def _prepare(self, prec):
        D = self.d_model
        q, s, c = self.qst_attn, self.slf_attn, self.crs_attn
        in_src_w = torch.cat([q.in_proj_weight[:D], s.in_proj_weight, c.in_proj_weight], 0)
        in_src_b = torch.cat([q.in_proj_bias[:D], s.in_proj_bias, c.in_proj_bias], 0)
        out_w = torch.cat([q.out_proj.weight, s.out_proj.weight, c.out_proj.weight], 1)
        out_b = q.out_proj.bias + s.out_proj.bias + c.out_proj.bias
        return dict(
            in_src=cabi.Lin(in_src_w, in_src_b, prec),
            in_words=cabi.Lin(q.in_proj_weight[D:], q.in_proj_bias[D:], prec),
            out_cat=cabi.Lin(out_w, out_b, prec),
            ffn1=cabi.Lin(self.linear1.weight, self.linear1.bias, prec),
            ffn2=cabi.Lin(self.linear2.weight, self.linear2.bias, prec),
            norm1=cabi.LN(self.norm1), norm2=cabi.LN(self.norm2))

    def fill_weights(self, a, prec: int):
        """weights + hyper-parameters of this block into a QatAvqArgs (its own call, or the one embedded in QatForwardArgs)"""
        w = self._prepared(prec)
        a.D, a.H = self.d_model, self.nhead
        a.in_src, a.in_words, a.out_cat, a.ffn1, a.ffn2 = (w[k].c for k in ("in_src", "in_words", "out_cat", "ffn1", "ffn2"))
        a.norm1, a.norm2 = w["norm1"].c, w["norm2"].c

    def forward(self, src_q: Tensor, src_v: Tensor, query: Tensor, visualize: bool = False):
        self._guard()
        prec = self._prec()
        w = self._prepared(prec)
        B, T, D = src_q.shape
        L = query.shape[1]
        cabi.require_cuda_f32(src_q, "src_q", (B, T, D))
        cabi.require_cuda_f32(src_v, "src_v", (B, T, D))
        cabi.require_cuda_f32(query, "query", (B, L, D))
        out_a, out_v = torch.empty_like(src_q), torch.empty_like(src_v)
        aw = vw = None
        if visualize:
            aw = torch.zeros(B, T, L, device=src_q.device, dtype=torch.float32)
            vw = torch.zeros(B, T, L, device=src_q.device, dtype=torch.float32)
        a = cabi.AvqArgs()
        a.B, a.T, a.L, a.prec = B, T, L, prec
        self.fill_weights(a, prec)
        a.audio, a.video, a.words = src_q.data_ptr(), src_v.data_ptr(), query.data_ptr()
        a.out_audio, a.out_video = out_a.data_ptr(), out_v.data_ptr()
        a.attw_audio, a.attw_video = cabi.ptr(aw), cabi.ptr(vw)
        cabi.run(cabi.OP_AVQ, "qat_avq_cross_attn_fwd", a, src_q.device)
        if visualize:
            return out_a, out_v, [aw, vw]
        return out_a, out_v


# ------------------------------------------------------------------------------------------------
class QstGrounding(_QatModule):
    """modules.py:268-347: question attends over a (length-2) list of global features."""

    def __init__(self, d_model: int = 512, nhead: int = 8, dropout: float = 0.1):
        super().__init__()
        self.act = nn.ReLU()
        self.norm = nn.LayerNorm(d_model)
        self.attn = nn.MultiheadAttention(d_model, nhead, dropout=0.1)
        self.mlp = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(), nn.Linear(d_model // 2, d_model))
        self.dropout = nn.Dropout(dropout)
        self.mlp.apply(_init_linear)
        self.d_model, self.nhead = d_model, nhead

    def _prepare(self, prec):
        D = self.d_model
        at = self.attn
        return dict(
            attn_q=cabi.Lin(at.in_proj_weight[:D], at.in_proj_bias[:D], prec),
            attn_kv=cabi.Lin(at.in_proj_weight[D:], at.in_proj_bias[D:], prec),
            attn_out=cabi.Lin(at.out_proj.weight, at.out_proj.bias, prec),
            mlp0=cabi.Lin(self.mlp[0].weight, self.mlp[0].bias, prec),
            mlp2=cabi.Lin(self.mlp[2].weight, self.mlp[2].bias, prec),
            norm=cabi.LN(self.norm))

    def fill_weights(self, a, prec: int, head: Optional[nn.Linear] = None):
        """weights into a QatGroundArgs; returns the object that keeps the head's prepared weights alive (or None)"""
        w = self._prepared(prec)
        a.D, a.H = self.d_model, self.nhead
        a.attn_q, a.attn_kv, a.attn_out, a.mlp0, a.mlp2 = (w[k].c for k in ("attn_q", "attn_kv", "attn_out", "mlp0", "mlp2"))
        a.norm = w["norm"].c
        if head is None:
            return None
        key = (cabi._prepared_epoch, head.weight.data_ptr(), head.weight._version, head.bias.data_ptr(), head.bias._version)
        cached = self.__dict__.get("_head_cache")
        if cached is None or cached[0] != key:
            cached = (key, cabi.Lin(head.weight, head.bias, cabi.PREC_FP32))
            self.__dict__["_head_cache"] = cached
        a.has_head, a.head = 1, cached[1].c
        return cached[1]

    def forward(self, qst: Tensor, data: Union[Tensor, List[Tensor]], head: Optional[nn.Linear] = None):
        """`head` (not in the reference signature) fuses net.py:141-142 `head(relu(.))` into the same call
        and returns (feat, logits)."""
        self._guard()
        prec = self._prec()
        w = self._prepared(prec)
        if isinstance(data, (list, tuple)):
            parts = [d.reshape(d.shape[0], -1, d.shape[-1]) for d in data]
            data = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        B, S, D = data.shape
        qst = qst.reshape(B, D)
        cabi.require_cuda_f32(qst, "qst", (B, D))
        out = torch.empty(B, D, device=qst.device, dtype=torch.float32)
        a = cabi.GroundArgs()
        a.B, a.prec = B, prec
        hold = self.fill_weights(a, prec, head)
        if S == 2:                                  # both call sites of the forward (net.py:136, 138): the fused length-2 form
            d0, d1 = data[:, 0].contiguous(), data[:, 1].contiguous()
            cabi.require_cuda_f32(d0, "data[0]", (B, D))
            a.qst, a.d0, a.d1 = qst.data_ptr(), d0.data_ptr(), d1.data_ptr()
        else:                                       # any other concatenated sequence (modules.py:318-326)
            data = data.contiguous()
            cabi.require_cuda_f32(data, "data", (B, S, D))
            a.qst, a.data, a.S = qst.data_ptr(), data.data_ptr(), S
        a.out = out.data_ptr()
        logits = None
        if head is not None:
            logits = torch.empty(B, head.weight.shape[0], device=qst.device, dtype=torch.float32)
            a.logits = logits.data_ptr()
        cabi.run(cabi.OP_GROUND, "qat_qst_grounding_head_fwd", a, qst.device)
        return (out, logits) if head is not None else out


# ------------------------------------------------------------------------------------------------
class TempMoE(_QatModule):
    """modules.py:350-612: question-aware Gaussian mixture of temporal experts."""

    def __init__(self, d_model: int = 512, nhead: int = 8, topK: int = 5, n_experts: int = 10, sigma: int = 9,
                 dropout: float = 0.1, vis_branch: bool = False):
        super().__init__()
        self.sigma, self.topK, self.n_experts = sigma, topK, n_experts
        if vis_branch:
            self.anorm = nn.LayerNorm(d_model)
            self.vnorm = nn.LayerNorm(d_model)
        else:
            self.norm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.qst_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.1)
        self.gauss_pred = nn.Sequential(nn.Linear(d_model, 2 * n_experts))
        self.router = nn.Sequential(nn.Linear(d_model, n_experts))
        self.experts = nn.ModuleList([
            nn.Sequential(nn.Linear(d_model, int(d_model // 2)), nn.ReLU(), nn.Linear(int(d_model // 2), d_model))
            for _ in range(n_experts)])
        self.experts.apply(_init_linear)
        self.margin = 1 / (n_experts * 2)
        self.center = torch.linspace(self.margin, 1 - self.margin, self.n_experts)   # plain attribute, as upstream
        self.center.requires_grad_(False)
        self.router.apply(_init_linear)
        self.gauss_pred.apply(_init_linear)
        self.d_model, self.nhead, self.vis_branch = d_model, nhead, vis_branch
        self.last_aux = None       # (topk_inds, topk_probs, gauss_weight) of the last call when `keep_aux`
        self.keep_aux = False

    def _prepare(self, prec):
        D, E = self.d_model, self.n_experts
        Ch = D // 2
        at = self.qst_attn
        # [router | gauss_pred] o out_proj o value projection, composed in fp64 (one query row per sample: the projections commute
        # with the attention, csrc/moe_front.cuh): attn_rg.w[:, h*D:(h+1)*D] = (Wrg Wo)[:, head h] Wv[head h, :]
        H, hd = self.nhead, D // self.nhead
        rg_w = torch.cat([self.router[0].weight, self.gauss_pred[0].weight], 0).double()
        rg_b = torch.cat([self.router[0].bias, self.gauss_pred[0].bias], 0).double()
        wo, bo = at.out_proj.weight.double(), at.out_proj.bias.double()
        wv, bv = at.in_proj_weight[2 * D:].double(), at.in_proj_bias[2 * D:].double()
        rgo = rg_w @ wo                                                                 # [3E, D]
        arg_w = torch.cat([rgo[:, h * hd:(h + 1) * hd] @ wv[h * hd:(h + 1) * hd] for h in range(H)], 1).float()   # [3E, H*D]
        arg_b = (rg_w @ (wo @ bv + bo) + rg_b).float()
        w1 = torch.cat([e[0].weight for e in self.experts], 0)
        b1 = torch.cat([e[0].bias for e in self.experts], 0)
        w2 = torch.zeros(D, E * Ch + 64, device=w1.device, dtype=torch.float32)
        for i, e in enumerate(self.experts):
            w2[:, i * Ch:(i + 1) * Ch] = e[2].weight
            w2[:, E * Ch + i] = e[2].bias
        d = dict(
            attn_q=cabi.Lin(at.in_proj_weight[:D], at.in_proj_bias[:D], prec),
            attn_kv=cabi.Lin(at.in_proj_weight[D:], at.in_proj_bias[D:], cabi.PREC_FP32),
            attn_rg=cabi.Lin(arg_w, arg_b, cabi.PREC_FP32),
            w1_cat=cabi.Lin(w1, b1, prec),
            w2_ext=cabi.Lin(w2, None, prec))
        if self.vis_branch:
            d["norm_a"], d["norm_v"] = cabi.LN(self.anorm), cabi.LN(self.vnorm)
        else:
            d["norm_a"] = d["norm_v"] = cabi.LN(self.norm)
        return d

    def fill_weights(self, a, prec: int):
        w = self._prepared(prec)
        a.D, a.H, a.E, a.K, a.sigma = self.d_model, self.nhead, self.n_experts, self.topK, float(self.sigma)
        a.attn_q, a.attn_kv, a.attn_rg, a.w1_cat, a.w2_ext = (w[k].c for k in ("attn_q", "attn_kv", "attn_rg", "w1_cat", "w2_ext"))
        a.norm_a, a.norm_v = w["norm_a"].c, w["norm_v"].c

    def attach_aux(self, a, B: int, T: int, dev):
        """optional rank / gate / mask outputs (`keep_aux`): allocate them and point the args at them"""
        if not self.keep_aux:
            return None
        aux = (torch.empty(B, self.topK, device=dev, dtype=torch.int64),
               torch.empty(B, self.topK, device=dev, dtype=torch.float32),
               torch.empty(B, self.topK, T, device=dev, dtype=torch.float32))
        a.topk_inds, a.topk_probs, a.gauss_weight = (t.data_ptr() for t in aux)
        return aux

    def forward(self, qst: Tensor, data: Tensor, sub_data: Optional[List[Tensor]] = None):
        self._guard()
        prec = self._prec()
        B, T, D = data.shape
        qst = qst.reshape(B, D)
        cabi.require_cuda_f32(qst, "qst", (B, D))
        cabi.require_cuda_f32(data, "data", (B, T, D))
        dev = data.device
        out_a = torch.empty(B, 1, D, device=dev, dtype=torch.float32)
        out_v = None
        a = cabi.MoeArgs()
        a.B, a.T, a.prec = B, T, prec
        self.fill_weights(a, prec)
        a.qst, a.data = qst.data_ptr(), data.data_ptr()
        if sub_data is not None:
            cabi.require_cuda_f32(sub_data[0], "sub_data[0]", (B, T, D))
            cabi.require_cuda_f32(sub_data[1], "sub_data[1]", (B, T, D))
            out_v = torch.empty(B, 1, D, device=dev, dtype=torch.float32)
            a.sub_a, a.sub_v = sub_data[0].data_ptr(), sub_data[1].data_ptr()
        a.out_a, a.out_v = out_a.data_ptr(), cabi.ptr(out_v)
        aux = self.attach_aux(a, B, T, dev)
        cabi.run(cabi.OP_MOE, "qat_temp_moe_fwd", a, dev)
        self.last_aux = aux
        if sub_data is not None:
            return out_a, out_v
        return out_a


def gauss_reduce(experts_logits: Tensor, router_logits: Tensor, gauss_raw: Tensor, topK: int, sigma: float = 9.0,
                 t_split: int = 0, return_aux: bool = False):
    """Standalone M2-M4 (`qat_gauss_reduce_fwd`): experts_logits [T,B,E,C] (fp32 or bf16, reference layout of
    modules.py:590 `torch.stack(dim=2)`), router_logits [B,E] (pre-softmax), gauss_raw [B,2E] -> [B,C]."""
    T, B, E, Cc = experts_logits.shape
    if not experts_logits.is_cuda or not experts_logits.is_contiguous():
        raise RuntimeError("gauss_reduce: experts_logits must be a contiguous CUDA tensor")
    if experts_logits.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("gauss_reduce: experts_logits must be float32 or bfloat16")
    cabi.require_cuda_f32(router_logits, "router_logits", (B, E))
    cabi.require_cuda_f32(gauss_raw, "gauss_raw", (B, 2 * E))
    dev = experts_logits.device
    out = torch.empty(B, Cc, device=dev, dtype=torch.float32)
    a = cabi.GaussArgs()
    a.B, a.T, a.E, a.K, a.C = B, T, E, topK, Cc
    a.dtype = cabi.DT_F32 if experts_logits.dtype == torch.float32 else cabi.DT_BF16
    a.sigma, a.stream = float(sigma), cabi.stream_ptr(dev)
    a.experts, a.stride_t, a.stride_b = experts_logits.data_ptr(), B * E * Cc, E * Cc
    a.router_logits, a.gauss_raw, a.out = router_logits.data_ptr(), gauss_raw.data_ptr(), out.data_ptr()
    a.t_split = t_split
    aux = None
    if return_aux:
        aux = (torch.empty(B, topK, device=dev, dtype=torch.int64), torch.empty(B, topK, device=dev, dtype=torch.float32),
               torch.empty(B, topK, T, device=dev, dtype=torch.float32))
        a.topk_inds, a.topk_probs, a.gauss_weight = (t.data_ptr() for t in aux)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().qat_gauss_reduce_fwd(C.byref(a)), "qat_gauss_reduce_fwd")
    return (out, aux) if return_aux else out


# ------------------------------------------------------------------------------------------------
class PatchSelecter(_QatModule):
    """modules.py:615-716.  `forward(patch, audio, video, proj=None)`: passing the model's `patch_proj`
    as `proj` lets the raw [B,T,P,1024] patches go straight into the fused path (net.py:124 + :130)."""

    def __init__(self, d_model: int = 512, nhead: int = 8, dropout: float = 0.1):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        self.vnorm = nn.LayerNorm(d_model)
        self.anorm = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.slf_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.crs_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.mlp = nn.Sequential(nn.Linear(d_model, d_model // 2), nn.ReLU(), nn.Linear(d_model // 2, d_model))
        self.mlp.apply(_init_linear)
        self.chunk_frames = 0

    def _prepare(self, prec):
        D = self.d_model
        s, c = self.slf_attn, self.crs_attn
        return dict(
            slf_in=cabi.Lin(s.in_proj_weight, s.in_proj_bias, prec, f16=True),
            slf_out=cabi.Lin(s.out_proj.weight, s.out_proj.bias, prec, f16=True),
            crs_q=cabi.Lin(c.in_proj_weight[:D], c.in_proj_bias[:D], prec, f16=True),
            crs_kv=cabi.Lin(c.in_proj_weight[D:], c.in_proj_bias[D:], prec, f16=True),
            crs_out=cabi.Lin(c.out_proj.weight, c.out_proj.bias, prec, f16=True),
            mlp0=cabi.Lin(self.mlp[0].weight, self.mlp[0].bias, prec, f16=True),
            mlp2=cabi.Lin(self.mlp[2].weight, self.mlp[2].bias, prec, f16=True),
            anorm=cabi.LN(self.anorm), vnorm=cabi.LN(self.vnorm))

    def fill_weights(self, a, prec: int, proj: Optional[Projection], Cin: int):
        w = self._prepared(prec)
        a.D, a.H, a.Cin, a.chunk_frames = self.d_model, self.nhead, Cin, int(self.chunk_frames)
        if proj is not None:
            a.has_proj, a.proj = 1, proj._prepared(prec).c
        for k in ("slf_in", "slf_out", "crs_q", "crs_kv", "crs_out", "mlp0", "mlp2"):
            setattr(a, k, w[k].c)
        a.anorm, a.vnorm = w["anorm"].c, w["vnorm"].c

    def forward(self, patch: Tensor, audio: Tensor, video: Tensor, proj: Optional[Projection] = None) -> List[Tensor]:
        self._guard()
        prec = self._prec()
        B, T, P, Cin = patch.shape
        D = self.d_model
        cabi.require_cuda_f32(patch, "patch", allow_f16=True)      # fp16 = half-size feature store
        cabi.require_cuda_f32(audio, "audio", (B, T, D))
        cabi.require_cuda_f32(video, "video", (B, T, D))
        if proj is None and Cin != D:
            raise RuntimeError(f"PatchSelecter: patch last dim {Cin} != d_model {D} and no `proj` given")
        out_a = torch.empty(B, T, D, device=patch.device, dtype=torch.float32)
        out_v = torch.empty(B, T, D, device=patch.device, dtype=torch.float32)
        a = cabi.PatchArgs()
        a.B, a.T, a.P, a.prec = B, T, P, prec
        self.fill_weights(a, prec, proj, Cin)
        if patch.dtype == torch.float16:
            a.patch_f16 = patch.data_ptr()
        else:
            a.patch = patch.data_ptr()
        a.audio, a.video = audio.data_ptr(), video.data_ptr()
        a.out_a, a.out_v = out_a.data_ptr(), out_v.data_ptr()
        cabi.run(cabi.OP_PATCH, "qat_patch_select_fwd", a, patch.device)
        return [out_a, out_v]
