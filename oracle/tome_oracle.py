"""CPU oracle for the ToMe bipartite soft-matching merge -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/src/tome/merge.py:18-97 (bipartite_soft_matching), :210-225 (merge_wavg),
:228-241 (merge_source) and the 23-step schedule driven by src/tome/patch/timm.py:35-65,
src/tome/utils.py:94-97 and scripts/extract_ToMe/extract_tome14.py:162 (r=[25]*23, class token
protected).  Index outputs are explicit arrays instead of closures.

Parity pin: tests/golden/tome_*.npz written by oracle/make_golden.py from the real merge.py
(loaded by file path) -- the reference holds no golden vectors of its own ("parity unpinned"
by upstream tests; pinned here against reference outputs generated in-container).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

Tensor = torch.Tensor


def clamp_r(t: int, r: int, class_token: bool = True, distill_token: bool = False) -> int:
    """merge.py:36-44: at most half of the unprotected tokens can be removed."""
    prot = int(class_token) + int(distill_token)
    return max(0, min(r, (t - prot) // 2))


def match(metric: Tensor, r: int, class_token: bool = True, distill_token: bool = False):
    """metric [N,t,C] -> (unm_idx [N,ta-r], src_idx [N,r], dst_idx [N,r], node_max, node_idx) or None if r<=0.

    Tokens at even positions form set A (ta = ceil(t/2)), odd positions set B (tb = floor(t/2)).
    """
    N, t, C = metric.shape
    r = clamp_r(t, r, class_token, distill_token)
    if r <= 0:
        return None
    m = metric / metric.norm(dim=-1, keepdim=True)                  # :50
    a, b = m[:, 0::2, :], m[:, 1::2, :]                             # :51
    scores = a @ b.transpose(-1, -2)                                # :52  [N,ta,tb]
    if class_token:
        scores[:, 0, :] = -math.inf                                 # :55
    if distill_token:
        scores[:, :, 0] = -math.inf                                 # :57
    node_max, node_idx = scores.max(dim=-1)                         # :59 first max index on CPU
    edge_idx = node_max.argsort(dim=-1, descending=True)            # :60
    unm_idx = edge_idx[:, r:]                                       # :62
    src_idx = edge_idx[:, :r]                                       # :63
    dst_idx = torch.gather(node_idx, 1, src_idx)                    # :64
    if class_token:
        unm_idx = unm_idx.sort(dim=1).values                        # :66-68
    return unm_idx, src_idx, dst_idx, node_max, node_idx


def _cat(unm: Tensor, dst: Tensor, distill_token: bool) -> Tensor:
    """merge.py:77-80: with a distillation token the class and distillation tokens lead the merged sequence"""
    if distill_token:
        return torch.cat([unm[:, :1], dst[:, :1], unm[:, 1:], dst[:, 1:]], dim=1)
    return torch.cat([unm, dst], dim=1)


def merge_apply(x: Tensor, unm_idx: Tensor, src_idx: Tensor, dst_idx: Tensor, mode: str = "sum", distill_token: bool = False) -> Tensor:
    """merge closure (:70-80): out = [A[unm_idx], B scatter-reduced with A[src_idx] at dst_idx]."""
    N, t, C = x.shape
    A, Bs = x[:, 0::2, :], x[:, 1::2, :].clone()
    unm = torch.gather(A, 1, unm_idx[..., None].expand(-1, -1, C))
    src = torch.gather(A, 1, src_idx[..., None].expand(-1, -1, C))
    for n in range(N):                                              # explicit scatter (include_self)
        for k in range(src_idx.shape[1]):
            d = int(dst_idx[n, k])
            if mode == "sum":
                Bs[n, d] += src[n, k]
            elif mode == "amax":
                Bs[n, d] = torch.maximum(Bs[n, d], src[n, k])
            else:
                raise ValueError(mode)
    return _cat(unm, Bs, distill_token)


def merge_apply_fast(x, unm_idx, src_idx, dst_idx, mode="sum", distill_token=False):
    """Same result as merge_apply via index_add (used for the bigger cases; sum order may differ
    in the last bit when several sources hit one destination)."""
    N, t, C = x.shape
    A, Bs = x[:, 0::2, :], x[:, 1::2, :].clone()
    unm = torch.gather(A, 1, unm_idx[..., None].expand(-1, -1, C))
    src = torch.gather(A, 1, src_idx[..., None].expand(-1, -1, C))
    Bs = Bs.scatter_reduce(1, dst_idx[..., None].expand(-1, -1, C), src, reduce=mode)
    return _cat(unm, Bs, distill_token)


def merge_wavg(x: Tensor, size: Optional[Tensor], idx, fast: bool = True, distill_token: bool = False):
    """merge.py:210-225: size-weighted mean.  size [N,t,1]."""
    if size is None:
        size = torch.ones_like(x[..., 0, None])
    f = merge_apply_fast if fast else merge_apply
    xs = f(x * size, *idx, mode="sum", distill_token=distill_token)
    size = f(size, *idx, mode="sum", distill_token=distill_token)
    return xs / size, size


def group_ids_update(group: Tensor, t: int, idx, distill_token: bool = False) -> Tensor:
    """Integer form of merge_source (:228-241).  group [N,T0] holds, per ORIGINAL token, its current
    token position; after a merge step it holds the position in the merged sequence."""
    unm_idx, src_idx, dst_idx = idx
    N = group.shape[0]
    ta = (t + 1) // 2
    n_unm = unm_idx.shape[1]
    newpos = torch.empty(N, t, dtype=torch.int64)
    for n in range(N):
        pu, pd = torch.arange(n_unm), n_unm + torch.arange(t // 2)
        if distill_token:                                           # [unm_0, dst_0, unm_1.., dst_1..]
            pu = torch.where(pu > 0, pu + 1, pu)
            pd[0] = 1
        newpos[n, 1::2] = pd
        newpos[n, 2 * unm_idx[n]] = pu
        newpos[n, 2 * src_idx[n]] = pd[dst_idx[n]]
    return torch.gather(newpos, 1, group)


def default_schedule(t0: int = 577, r: int = 25, blocks: int = 24, steps: int = 23) -> List[int]:
    """extract_tome14.py:162 `model.r=[25]*23`, padded with 0 to 24 blocks (tome/utils.py:94-97)."""
    return [r] * steps + [0] * (blocks - steps)


def standin_metric(x: Tensor) -> Tensor:
    """the real metric (mean over heads of K inside the ViT block) is out of
    scope, so the chain is driven by x.view(N,t,16,64).mean(2)."""
    N, t, C = x.shape
    return x.view(N, t, C // 64, 64).mean(2)


def chain(x: Tensor, schedule: Optional[List[int]] = None, class_token: bool = True, fast: bool = True):
    """Run the whole schedule.  Returns (x [N,t_end,C], size [N,t_end,1], group [N,T0], per-step idx list)."""
    N, t0, C = x.shape
    if schedule is None:
        schedule = default_schedule(t0)
    size = None
    group = torch.arange(t0)[None, :].expand(N, t0).contiguous()
    steps = []
    for r in schedule:
        t = x.shape[1]
        res = match(standin_metric(x), r, class_token=class_token)
        if res is None:
            continue
        idx = res[:3]
        steps.append(idx)
        group = group_ids_update(group, t, idx)
        x, size = merge_wavg(x, size, idx, fast=fast)
    if size is None:
        size = torch.ones_like(x[..., 0, None])
    return x, size, group, steps


# ---- the ToMe-patched ViT block (tome/patch/timm.py) ----------------------------------------------------------
def tome_attention(sd, p: str, x: Tensor, size: Optional[Tensor], H: int):
    """ToMeAttention.forward, timm.py:75-108: x is the attention input (norm1 already applied by the block).  Returns
    (proj(attention) [N,t,C], metric = k.mean(heads) [N,t,C/H])."""
    N, t, C = x.shape
    hd = C // H
    qkv = (x @ sd[p + ".qkv.weight"].t() + sd[p + ".qkv.bias"]).reshape(N, t, 3, H, hd).permute(2, 0, 3, 1, 4)   # :79-83
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * hd ** -0.5                                                              # :90
    if size is not None:
        attn = attn + size.log()[:, None, None, :, 0]                                                         # :94-95
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(N, t, C)
    o = o @ sd[p + ".proj.weight"].t() + sd[p + ".proj.bias"]                                                  # :100
    return o, k.mean(1)                                                                                        # :107


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def tome_block(sd, p: str, x: Tensor, size: Optional[Tensor], r: int, H: int, class_token: bool = True, prop_attn: bool = True,
               eps: float = 1e-6, group: Optional[Tensor] = None):
    """ToMeBlock.forward, timm.py:35-65.  `p` prefixes timm's block parameter names.  Returns
    (x_out, size_out | None, metric, idx | None, group | None)."""
    N, t, C = x.shape
    xa, metric = tome_attention(sd, p + ".attn", _ln(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps), size if prop_attn else None, H)   # :39-40
    x = x + xa                                                                                                 # :41
    idx = None
    if r > 0:
        res = match(metric, r, class_token=class_token)                                                        # :48-53
        if res is not None:
            idx = res[:3]
            if group is not None:
                group = group_ids_update(group, t, idx)                                                        # :54-57 (as group ids)
            x, size = merge_wavg(x, size, idx)                                                                 # :58
    h = _ln(x, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps) @ sd[p + ".mlp.fc1.weight"].t() + sd[p + ".mlp.fc1.bias"]
    h = torch.nn.functional.gelu(h)
    x = x + (h @ sd[p + ".mlp.fc2.weight"].t() + sd[p + ".mlp.fc2.bias"])                                      # :62
    return x, size, metric, idx, group
