"""CPU oracle for the QA-TIGER eval hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file.  The product path (`qa-tiger_b200/`) never does; it fails loudly
when the CUDA extension is missing.

What this is: a plain functional restatement (torch CPU tensors used as an array library,
fp32) of the reference's eval forward.  All citations are relative to /root/reference/.

Parity pin: the reference ships NO golden vectors / tests for this path,
so the pin is "outputs of the reference itself run here": `oracle/make_golden.py` imports the
real `src/models/net.py` / `modules.py` / `tome/merge.py`, loads the seeded weights from
`oracle/synth.py`, and writes `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks this
restatement against those files (<= 2e-5 normwise on logits, identical indices).

Third-party arithmetic: `torch.nn.functional.multi_head_attention_forward`, `layer_norm`,
`topk`, `softmax` live in torch (reference pins torch==2.4.0 in pyproject.toml:26; not
vendored).  Their published algorithm is restated in `mha()` / `layer_norm()` below and is
anchored on the reference call sites modules.py:217-221,338,562,688,698.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# Precision-emulation hook (used only by the error-budget study in tests/DESIGN.md):
# a callable (x, w, tag) -> (x', w') applied to the operands of every dense contraction.
# ----------------------------------------------------------------------------------------
_OPERAND_HOOK = None


def set_operand_hook(fn):
    global _OPERAND_HOOK
    _OPERAND_HOOK = fn


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], tag: str = "") -> Tensor:
    """y = x W^T + b  (nn.Linear; modules.py:35-43 Projection and every Linear on the path)."""
    if _OPERAND_HOOK is not None:
        x, w = _OPERAND_HOOK(x, w, tag)
    y = x @ w.t()
    if b is not None:
        y = y + b
    return y


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """LayerNorm over the last axis, biased variance, eps=1e-5 (nn.LayerNorm default)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * g + b


def mha(xq: Tensor, xkv: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
        nhead: int = 8, tag: str = "", return_weights: bool = False, causal: bool = False):
    """nn.MultiheadAttention slow path on batch-first views.

    xq [N,Lq,D], xkv [N,Lk,D] (key is value at every reference call site).
    Order: in-proj (rows 0:D for q, D:3D for k|v) -> split heads -> q*hd^-0.5 -> QK^T ->
    softmax -> PV -> merge heads -> out-proj; optional head-averaged weights.
    No key padding mask anywhere (modules.py:217 passes none).
    """
    N, Lq, D = xq.shape
    Lk = xkv.shape[1]
    hd = D // nhead
    q = linear(xq, in_w[:D], in_b[:D], tag + ".q")
    k = linear(xkv, in_w[D:2 * D], in_b[D:2 * D], tag + ".k")
    v = linear(xkv, in_w[2 * D:], in_b[2 * D:], tag + ".v")
    q = q.view(N, Lq, nhead, hd).permute(0, 2, 1, 3) * (hd ** -0.5)
    k = k.view(N, Lk, nhead, hd).permute(0, 2, 1, 3)
    v = v.view(N, Lk, nhead, hd).permute(0, 2, 1, 3)
    s = q @ k.transpose(-1, -2)                              # [N,H,Lq,Lk]
    if causal:
        s = s + torch.full((Lq, Lk), float("-inf")).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(N, Lq, D)
    o = linear(o, out_w, out_b, tag + ".o")
    if return_weights:
        return o, p.mean(dim=1)                              # [N,Lq,Lk] head-averaged
    return o


def _mha_w(sd: Dict[str, Tensor], p: str):
    return (sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"],
            sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


# ----------------------------------------------------------------------------------------
# A1/A2  AVQCrossAttn  (modules.py:194-265)
# ----------------------------------------------------------------------------------------
def avq_sub_forward(sd, p, src_q, src_v, words, return_weights=False):
    """One direction (modules.py:194-237): LN1(src_q+slf+crs+qst) then LN2(x + W2 relu(W1 x))."""
    qst, w = mha(src_q, words, *_mha_w(sd, p + ".qst_attn"), tag=p + ".qst", return_weights=True)
    slf = mha(src_q, src_q, *_mha_w(sd, p + ".slf_attn"), tag=p + ".slf")
    crs = mha(src_q, src_v, *_mha_w(sd, p + ".crs_attn"), tag=p + ".crs")
    x = src_q + slf + crs + qst                                      # :223-226 (dropout = id)
    x = layer_norm(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    h = torch.relu(linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"], p + ".ffn1"))
    x = x + linear(h, sd[p + ".linear2.weight"], sd[p + ".linear2.bias"], p + ".ffn2")
    x = layer_norm(x, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    return (x, w) if return_weights else x


def avq_cross_attn(sd, p, audio, video, words, visualize=False):
    """modules.py:239-265.  The second direction consumes the ORIGINAL audio (:257-259)."""
    a, aw = avq_sub_forward(sd, p, audio, video, words, True)
    v, vw = avq_sub_forward(sd, p, video, audio, words, True)
    if visualize:
        return a, v, [aw, vw]
    return a, v


# ----------------------------------------------------------------------------------------
# S1  PatchSelecter  (modules.py:656-716)
# ----------------------------------------------------------------------------------------
def patch_selecter(sd, p, patch, audio, video):
    """patch [B,T,P,D], audio/video [B,T,D] -> [anorm(a), vnorm(v)] each [B,T,D]."""
    B, T, P, D = patch.shape
    pt = patch.reshape(B * T, P, D)
    pt = pt + mha(pt, pt, *_mha_w(sd, p + ".slf_attn"), tag=p + ".slf")        # :688 (no LN)
    q = torch.stack([video.reshape(B * T, D), audio.reshape(B * T, D)], dim=1)  # :692 video first
    o = mha(q, pt, *_mha_w(sd, p + ".crs_attn"), tag=p + ".crs")                # :698
    h = torch.relu(linear(o, sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"], p + ".mlp0"))
    o = linear(h, sd[p + ".mlp.2.weight"], sd[p + ".mlp.2.bias"], p + ".mlp2")  # :702
    v, a = o[:, 0].reshape(B, T, D), o[:, 1].reshape(B, T, D)                   # :706
    return [layer_norm(a, sd[p + ".anorm.weight"], sd[p + ".anorm.bias"]),      # :713-716
            layer_norm(v, sd[p + ".vnorm.weight"], sd[p + ".vnorm.bias"])]


# ----------------------------------------------------------------------------------------
# M2/M3  router + Gaussian masks   (modules.py:565-578, 420-471)
# ----------------------------------------------------------------------------------------
def router_topk(router_logits: Tensor, K: int) -> Tuple[Tensor, Tensor]:
    """softmax over ALL E, top-K (descending, ties -> lowest index as torch CPU), renormalise."""
    probs = torch.softmax(router_logits, dim=-1)
    B, E = probs.shape
    # explicit stable selection so the tie rule is spelled out rather than inherited
    inds = torch.empty(B, K, dtype=torch.int64)
    vals = torch.empty(B, K, dtype=probs.dtype)
    work = probs.clone()
    for i in range(K):
        m = work.max(dim=-1).values
        first = (work == m[:, None]).to(torch.int64).argmax(dim=-1)   # lowest index among ties
        inds[:, i] = first
        vals[:, i] = m
        work[torch.arange(B), first] = -1.0
    vals = vals / vals.sum(dim=-1, keepdim=True)
    return vals, inds


def gaussian_masks(gauss_raw: Tensor, topk_inds: Tensor, E: int, T: int, sigma: float = 9.0) -> Tensor:
    """gauss_raw [B,2E] (pre-activation gauss_pred output) -> peak-normalised masks [B,K,T].

    centre_e = linspace(m,1-m,E)[e] + tanh(g[e,0])*m, m = 1/(2E)   (modules.py:405-407,439,576)
    width_e  = sigmoid(g[e,1])                                      (:577)
    gather by topk_inds; c = clamp(centre,0,1); w = clamp(width,min=.09)/sigma   (:454-457)
    G = (0.3989422804014327/w) * exp(-(t/(T-1)-c)^2/(2 w^2)); G /= max_t G       (:460-469)
    """
    B = gauss_raw.shape[0]
    g = gauss_raw.view(B, E, 2)
    margin = 1.0 / (E * 2)
    base = torch.linspace(margin, 1 - margin, E)
    centre = base[None, :] + torch.tanh(g[:, :, 0]) * margin
    width = torch.sigmoid(g[:, :, 1])
    c = torch.gather(centre, 1, topk_inds).clamp(min=0, max=1)[:, :, None]       # [B,K,1]
    w = (torch.gather(width, 1, topk_inds).clamp(min=0.09) / sigma)[:, :, None]
    t = torch.linspace(0, 1, T)[None, None, :]
    G = 0.3989422804014327 / w * torch.exp(-(t - c) ** 2 / (2 * w ** 2))
    return G / G.max(dim=-1, keepdim=True).values


# ----------------------------------------------------------------------------------------
# M4  get_output, including the batch-coupling quirk   (modules.py:473-534)
# ----------------------------------------------------------------------------------------
def moe_combine(experts_logits: Tensor, gauss_weight: Tensor, topk_inds: Tensor,
                topk_probs: Tensor) -> Tensor:
    """experts_logits [T,B,E,C] -> [B,C].

    out[b] = sum_i p[b,i] * sum_t G[b,i,t] * X[t, b, j(b,t,i), :],
    j(b,t,i) = topk_inds[(b*T+t) % B, i]    <- `topk_inds.repeat(T,1)` (:502) against rows
    ordered b*T+t (:495): row (b,t) borrows the rank list of sample (b*T+t) mod B.
    """
    T, B, E, C = experts_logits.shape
    K = topk_inds.shape[1]
    rows = torch.arange(B * T)
    owner = rows % B                                              # which sample's rank list
    j = topk_inds[owner]                                          # [B*T,K]
    x = experts_logits.permute(1, 0, 2, 3).reshape(B * T, E, C)
    sel = x[rows[:, None], j]                                     # [B*T,K,C]
    sel = sel.view(B, T, K, C)
    wgt = (gauss_weight.permute(0, 2, 1) * topk_probs[:, None, :])   # [B,T,K]
    return (sel * wgt[..., None]).sum(dim=(1, 2))


def moe_coef(gauss_weight: Tensor, topk_inds: Tensor, topk_probs: Tensor, E: int) -> Tensor:
    """coef[b,t,j] = sum_{i: j(b,t,i)=j} p[b,i]*G[b,i,t]  (SURVEY 8a-M4 refactor) -> [B,T,E]."""
    B, K, T = gauss_weight.shape
    rows = torch.arange(B * T)
    j = topk_inds[rows % B].view(B, T, K)
    wgt = gauss_weight.permute(0, 2, 1) * topk_probs[:, None, :]
    coef = torch.zeros(B, T, E, dtype=gauss_weight.dtype)
    coef.scatter_add_(2, j, wgt)
    return coef


def gauss_reduce(experts_logits: Tensor, router_logits: Tensor, gauss_raw: Tensor, K: int,
                 sigma: float = 9.0):
    """Standalone M2-M4 (what `qat_gauss_reduce_fwd` computes).  Returns (out[B,C], aux)."""
    T, B, E, C = experts_logits.shape
    p, idx = router_topk(router_logits, K)
    G = gaussian_masks(gauss_raw, idx, E, T, sigma)
    out = moe_combine(experts_logits, G, idx, p)
    return out, dict(topk_probs=p, topk_inds=idx, gauss_weight=G)


# ----------------------------------------------------------------------------------------
# M1  TempMoE.forward  (modules.py:536-612)
# ----------------------------------------------------------------------------------------
def _experts(sd, p, x: Tensor, E: int) -> Tensor:
    """x [B,T,D] -> experts_logits [T,B,E,C]  (:590/598/607 `torch.stack(dim=2)` on seq-first)."""
    outs = []
    for e in range(E):
        h = torch.relu(linear(x, sd[f"{p}.experts.{e}.0.weight"], sd[f"{p}.experts.{e}.0.bias"], p + ".e1"))
        outs.append(linear(h, sd[f"{p}.experts.{e}.2.weight"], sd[f"{p}.experts.{e}.2.bias"], p + ".e2"))
    return torch.stack(outs, dim=2).permute(1, 0, 2, 3)


def temp_moe(sd, p, qst, data, sub_data=None, K=7, E=7, sigma=9.0, aux: Optional[dict] = None):
    """qst [B,D], data [B,T,D], sub_data = [a,v] each [B,T,D] or None.

    Router / Gaussians come from `data` alone (:562) and are shared by both branches; experts see
    data+sub_data[i] (:587,597).  Returns [B,1,D] or a pair.
    """
    B, T, D = data.shape
    temp_w = mha(qst[:, None, :], data, *_mha_w(sd, p + ".qst_attn"), tag=p + ".qst")[:, 0]
    rl = linear(temp_w, sd[p + ".router.0.weight"], sd[p + ".router.0.bias"], "fp32.router")
    gr = linear(temp_w, sd[p + ".gauss_pred.0.weight"], sd[p + ".gauss_pred.0.bias"], "fp32.gauss")
    probs, inds = router_topk(rl, K)
    G = gaussian_masks(gr, inds, E, T, sigma)
    if aux is not None:
        aux.update(temp_w=temp_w, router_logits=rl, gauss_raw=gr, topk_probs=probs, topk_inds=inds,
                   gauss_weight=G)
    if sub_data is not None:
        outs = []
        for i, nm in enumerate(("anorm", "vnorm")):
            x = _experts(sd, p, data + sub_data[i], E)
            o = moe_combine(x, G, inds, probs)[:, None, :]
            outs.append(layer_norm(o, sd[f"{p}.{nm}.weight"], sd[f"{p}.{nm}.bias"]))
        return outs[0], outs[1]
    x = _experts(sd, p, data, E)
    o = moe_combine(x, G, inds, probs)[:, None, :]
    return layer_norm(o, sd[p + ".norm.weight"], sd[p + ".norm.bias"])


# ----------------------------------------------------------------------------------------
# G1  QstGrounding  (modules.py:304-347)
# ----------------------------------------------------------------------------------------
def qst_grounding(sd, p, qst, data):
    """qst [B,D]; data = Tensor [B,S,D] or list of them (concatenated on the sequence axis)."""
    if isinstance(data, (list, tuple)):
        data = torch.cat(list(data), dim=1)
    a = mha(qst[:, None, :], data, *_mha_w(sd, p + ".attn"), tag=p + ".attn")[:, 0]
    h = torch.relu(linear(a, sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"], p + ".mlp0"))
    f = data.mean(dim=1) + linear(h, sd[p + ".mlp.2.weight"], sd[p + ".mlp.2.bias"], p + ".mlp2")
    return layer_norm(f, sd[p + ".norm.weight"], sd[p + ".norm.bias"])


# ----------------------------------------------------------------------------------------
# Frozen CLIP text tower  (encoders.py:30-37, base/clip_base.py:186-254,390-396)
# out of kernel scope, restated so the full forward can be checked where the reference is absent
# ----------------------------------------------------------------------------------------
def clip_text_encoder(sd, p, text: Tensor, layers: int = 12, heads: int = 12):
    """text int[B,77] -> (quest [B,768] = ln_final(x)[eot] @ text_projection, words [B,77,768])."""
    x = sd[p + ".token_embedding.weight"][text.long()] + sd[p + ".positional_embedding"][: text.shape[1]]
    for l in range(layers):
        q = f"{p}.transformer.resblocks.{l}"
        y = layer_norm(x, sd[q + ".ln_1.weight"], sd[q + ".ln_1.bias"])
        x = x + mha(y, y, *_mha_w(sd, q + ".attn"), nhead=heads, tag="text.attn", causal=True)
        y = layer_norm(x, sd[q + ".ln_2.weight"], sd[q + ".ln_2.bias"])
        y = linear(y, sd[q + ".mlp.c_fc.weight"], sd[q + ".mlp.c_fc.bias"], "text.fc")
        y = y * torch.sigmoid(1.702 * y)                            # QuickGELU clip_base.py:162-164
        x = x + linear(y, sd[q + ".mlp.c_proj.weight"], sd[q + ".mlp.c_proj.bias"], "text.proj")
    x = layer_norm(x, sd[p + ".ln_final.weight"], sd[p + ".ln_final.bias"])
    eot = text.long().argmax(dim=-1)
    quest = x[torch.arange(x.shape[0]), eot] @ sd[p + ".text_projection"]
    return quest, x


# ----------------------------------------------------------------------------------------
# QA_TIGER.forward  (net.py:106-144)
# ----------------------------------------------------------------------------------------
def qa_tiger_forward(sd, batch: Dict[str, Tensor], K=7, E=7, text_feats=None, with_aux=False):
    """batch: quest int[B,77], audio [B,T,128], video [B,T,768], patch [B,T,P,1024] -> logits [B,42].

    text_feats=(quest768, words768) bypasses the text tower (used to time the hot path alone).
    """
    aux = {}
    if text_feats is None:
        quest, words = clip_text_encoder(sd, "quest_encoder", batch["quest"])
    else:
        quest, words = text_feats
    aux["quest768"], aux["words768"] = quest, words
    lin = lambda n, x: linear(x, sd[n + ".proj.weight"], sd[n + ".proj.bias"], n)
    audio = lin("audio_proj", batch["audio"])
    video = lin("video_proj", batch["video"])
    words = lin("words_proj", words)
    quest = lin("quest_proj", quest)
    patch = lin("patch_proj", batch["patch"])
    audio, video = avq_cross_attn(sd, "crs_attn", audio, video, words)
    aux["audio_x"], aux["video_x"] = audio, video
    a_p, v_p = patch_selecter(sd, "patch_selecter", patch, audio, video)
    aux["a_patch"], aux["v_patch"] = a_p, v_p
    at_aux, vt_aux = {}, {}
    a_global = temp_moe(sd, "at_aggregator", quest, audio, None, K, E, aux=at_aux)
    ap_global, vp_global = temp_moe(sd, "vt_aggregator", quest, video, [a_p, v_p], K, E, aux=vt_aux)
    aux["at"], aux["vt"] = at_aux, vt_aux
    aux["a_global"], aux["ap_global"], aux["vp_global"] = a_global, ap_global, vp_global
    fusion = qst_grounding(sd, "quest_grounding", quest, [ap_global, vp_global])
    aux["fusion1"] = fusion
    fusion = qst_grounding(sd, "quest_grounding", quest, [fusion[:, None, :], a_global])
    aux["fusion2"] = fusion
    out = linear(torch.relu(fusion), sd["head.weight"], sd["head.bias"], "fp32.head")   # :141-142
    return (out, aux) if with_aux else out
