"""Drop-in mirror of /root/reference/src/tome/patch/timm.py: the ToMe-patched ViT block on the B200 path.

  ToMeAttention.forward(x, size) -> (x_attn, metric)     timm.py:68-108  (proportional attention :94-95, metric = k.mean(heads) :107)
  ToMeBlock.forward(x)           -> x                    timm.py:22-65   (attention, bipartite_soft_matching on the metric,
                                                                          merge_source, merge_wavg, mlp)
  parse_r                                                 tome/utils.py:80-105
  ToMeStack                                               the r / size / source bookkeeping of ToMeVisionTransformer.forward :118-125

Parameter names are timm's (`norm1`, `attn.qkv`, `attn.proj`, `norm2`, `mlp.fc1`, `mlp.fc2`), so a timm ViT block's state_dict
loads unchanged.  The forward bodies contain no math: ToMeBlock.forward is ONE C-ABI call (`qat_tome_block_fwd`), ToMeAttention
alone is `qat_tome_attention_fwd`.  Eval only, CUDA only, no fallback.  Sources are tracked as group ids (position of every
original token in the current token list), from which `source_matrix()` expands the reference's one-hot adjacency on request.
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import _cabi as cabi
from ..models.modules import _QatModule


def parse_r(num_layers: int, r: Union[List[int], Tuple[int, float], int]) -> List[int]:
    """tome/utils.py:80-105: a constant r, an (r, inflection) pair or an explicit list -> per-layer list."""
    inflect = 0
    if isinstance(r, list):
        if len(r) < num_layers:
            r = r + [0] * (num_layers - len(r))
        return list(r)
    if isinstance(r, tuple):
        r, inflect = r
    min_val = int(r * (1.0 - inflect))
    max_val = 2 * r - min_val
    step = (max_val - min_val) / max(num_layers - 1, 1)        # (the reference divides by zero for a single layer)
    return [int(min_val + step * i) for i in range(num_layers)]


class ToMeAttention(_QatModule):
    """timm Attention parameter names; forward(x, size=None) -> (proj(attention), k.mean(heads))."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def _prepare(self, prec):
        return dict(qkv=cabi.Lin(self.qkv.weight, self.qkv.bias, prec), proj=cabi.Lin(self.proj.weight, self.proj.bias, prec))

    def forward(self, x: torch.Tensor, size: Optional[torch.Tensor] = None):
        self._guard()
        prec = self._prec()
        w = self._prepared(prec)
        N, t, Cc = x.shape
        cabi.require_cuda_f32(x, "x", (N, t, Cc))
        a = cabi.TomeBlockArgs()
        a.N, a.t, a.C, a.H, a.r, a.class_token, a.prop_attn, a.prec, a.eps = N, t, Cc, self.num_heads, 0, 0, int(size is not None), prec, 1e-6
        a.x = x.data_ptr()
        sz = None
        if size is not None:
            sz = size.reshape(N, t).contiguous().float()
            a.size = sz.data_ptr()
        a.qkv, a.proj = w["qkv"].c, w["proj"].c
        out = torch.empty_like(x)
        metric = torch.empty(N, t, 64, device=x.device, dtype=torch.float32)
        a.x_out, a.metric = out.data_ptr(), metric.data_ptr()
        cabi.run(cabi.OP_TOME_ATTENTION, "qat_tome_attention_fwd", a, x.device)
        return out, metric


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class ToMeBlock(_QatModule):
    """timm Block parameter names.  `_tome_info` is the dict the reference shares between the blocks of a model
    (r list, size, source, class_token, prop_attn, trace_source); `source` holds int32 group ids [N, T0] here."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = ToMeAttention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.eps = eps
        self._tome_info = dict(r=[0], size=None, source=None, trace_source=False, prop_attn=True, class_token=True, distill_token=False)
        self.last_indices = None      # (unm_idx, src_idx, dst_idx) int32 of the last forward when `keep_indices`
        self.keep_indices = False
        self.last_metric = None

    def _prepare(self, prec):
        return dict(qkv=cabi.Lin(self.attn.qkv.weight, self.attn.qkv.bias, prec), proj=cabi.Lin(self.attn.proj.weight, self.attn.proj.bias, prec),
                    fc1=cabi.Lin(self.mlp.fc1.weight, self.mlp.fc1.bias, prec), fc2=cabi.Lin(self.mlp.fc2.weight, self.mlp.fc2.bias, prec),
                    norm1=cabi.LN(self.norm1), norm2=cabi.LN(self.norm2))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._guard()
        info = self._tome_info
        flags = int(bool(info["class_token"])) | (2 if info.get("distill_token") else 0)     # the C ABI's protected-token flag word
        prec = self._prec()
        w = self._prepared(prec)
        N, t, Cc = x.shape
        cabi.require_cuda_f32(x, "x", (N, t, Cc))
        r = info["r"].pop(0)
        rr = cabi.lib().qat_tome_clamp_r(t, r, flags) if r > 0 else 0
        a = cabi.TomeBlockArgs()
        a.N, a.t, a.C, a.H, a.r, a.class_token, a.prop_attn, a.prec, a.eps = (N, t, Cc, self.attn.num_heads, r, flags,
                                                                           int(bool(info["prop_attn"])), prec, self.eps)
        a.x = x.data_ptr()
        size = info["size"]
        if size is not None:
            size = size.reshape(N, t).contiguous()
            a.size = size.data_ptr()
        for k in ("qkv", "proj", "fc1", "fc2", "norm1", "norm2"):
            setattr(a, k, w[k].c)
        out = torch.empty(N, t - rr, Cc, device=x.device, dtype=torch.float32)
        size_out = torch.empty(N, t - rr, device=x.device, dtype=torch.float32)
        a.x_out, a.size_out = out.data_ptr(), size_out.data_ptr()
        metric = idx = None
        if self.keep_indices:
            metric = torch.empty(N, t, 64, device=x.device, dtype=torch.float32)
            a.metric = metric.data_ptr()
            if rr > 0:
                ta = (t + 1) // 2
                idx = (torch.empty(N, ta - rr, device=x.device, dtype=torch.int32), torch.empty(N, rr, device=x.device, dtype=torch.int32),
                       torch.empty(N, rr, device=x.device, dtype=torch.int32))
                a.unm_idx, a.src_idx, a.dst_idx = (i.data_ptr() for i in idx)
        if info.get("trace_source") and rr > 0:
            if info["source"] is None:
                info["source"] = torch.arange(t, device=x.device, dtype=torch.int32).repeat(N, 1).contiguous()
            a.group, a.T0 = info["source"].data_ptr(), info["source"].shape[1]
        cabi.run(cabi.OP_TOME_BLOCK, "qat_tome_block_fwd", a, x.device)
        if rr > 0 or size is not None:
            info["size"] = size_out[..., None]
        self.last_indices, self.last_metric = idx, metric
        return out


def source_matrix(group: torch.Tensor, t_now: int) -> torch.Tensor:
    """group ids [N, T0] -> the reference's adjacency `source` [N, t_now, T0] (merge_source, merge.py:228-241)."""
    N, T0 = group.shape
    src = torch.zeros(N, t_now, T0, device=group.device, dtype=torch.float32)
    src.scatter_(1, group.long()[:, None, :], 1.0)
    return src


class ToMeStack(nn.Module):
    """A list of ToMeBlocks run with one r schedule: the bookkeeping of ToMeVisionTransformer.forward (timm.py:118-125)."""

    def __init__(self, blocks: List[ToMeBlock], r: Union[List[int], Tuple[int, float], int] = 0, trace_source: bool = False,
                 prop_attn: bool = True, class_token: bool = True, distill_token: bool = False):
        super().__init__()
        self.blocks = nn.ModuleList(blocks)
        self.r = r
        self._tome_info = dict(r=[0], size=None, source=None, trace_source=trace_source, prop_attn=prop_attn, class_token=class_token,
                               distill_token=distill_token)
        for b in self.blocks:
            b._tome_info = self._tome_info

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._tome_info["r"] = parse_r(len(self.blocks), self.r)
        self._tome_info["size"] = None
        self._tome_info["source"] = None
        for b in self.blocks:
            x = b(x)
        return x
