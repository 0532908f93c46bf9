"""Seeded synthetic weights and MUSIC-AVQA-shaped batches -- TEST/BENCH INFRASTRUCTURE.

Shared by oracle/make_golden.py (which loads these weights INTO the real reference model),
tests/ and bench.py, so that the reference, the oracle and the CUDA path all see identical
numbers without /root/reference being present at run time.

Key names / shapes follow the reference state_dict (net.py:34-61,
modules.py:170-182,283-292,366-403,630-646; encoders.py:14-22).  Values are NOT the
reference's default init: every bias and LayerNorm affine is randomised so the bias / affine
code paths are exercised (default init would zero them).
All generation is on CPU with an explicit torch.Generator => identical on every box of this image.
"""
from __future__ import annotations

from typing import Dict

import torch


def _mha_keys(sd, g, p, d):
    sd[p + ".in_proj_weight"] = torch.randn(3 * d, d, generator=g) * (1.0 / d ** 0.5)
    sd[p + ".in_proj_bias"] = torch.randn(3 * d, generator=g) * 0.05
    sd[p + ".out_proj.weight"] = torch.randn(d, d, generator=g) * (1.0 / d ** 0.5)
    sd[p + ".out_proj.bias"] = torch.randn(d, generator=g) * 0.05


def _lin(sd, g, p, out_f, in_f, gain=2.0):
    sd[p + ".weight"] = torch.randn(out_f, in_f, generator=g) * (gain / in_f) ** 0.5
    sd[p + ".bias"] = torch.randn(out_f, generator=g) * 0.05


def _ln(sd, g, p, d):
    sd[p + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
    sd[p + ".bias"] = 0.1 * torch.randn(d, generator=g)


def make_state_dict(seed: int = 713, d_model: int = 512, video_dim: int = 768, patch_dim: int = 1024,
                    audio_dim: int = 128, num_experts: int = 7, text_tower: bool = True,
                    text_width: int = 768, text_layers: int = 12) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D = d_model
    _lin(sd, g, "audio_proj.proj", D, audio_dim)
    _lin(sd, g, "video_proj.proj", D, video_dim)
    _lin(sd, g, "patch_proj.proj", D, patch_dim)
    _lin(sd, g, "words_proj.proj", D, video_dim)
    _lin(sd, g, "quest_proj.proj", D, video_dim)
    for a in ("qst_attn", "crs_attn", "slf_attn"):
        _mha_keys(sd, g, "crs_attn." + a, D)
    _lin(sd, g, "crs_attn.linear1", D, D)
    _lin(sd, g, "crs_attn.linear2", D, D)
    _ln(sd, g, "crs_attn.norm1", D)
    _ln(sd, g, "crs_attn.norm2", D)
    _ln(sd, g, "patch_selecter.vnorm", D)
    _ln(sd, g, "patch_selecter.anorm", D)
    _mha_keys(sd, g, "patch_selecter.slf_attn", D)
    _mha_keys(sd, g, "patch_selecter.crs_attn", D)
    _lin(sd, g, "patch_selecter.mlp.0", D // 2, D)
    _lin(sd, g, "patch_selecter.mlp.2", D, D // 2)
    _ln(sd, g, "quest_grounding.norm", D)
    _mha_keys(sd, g, "quest_grounding.attn", D)
    _lin(sd, g, "quest_grounding.mlp.0", D // 2, D)
    _lin(sd, g, "quest_grounding.mlp.2", D, D // 2)
    for agg, norms in (("at_aggregator", ("norm",)), ("vt_aggregator", ("anorm", "vnorm"))):
        for n in norms:
            _ln(sd, g, f"{agg}.{n}", D)
        _mha_keys(sd, g, f"{agg}.qst_attn", D)
        _lin(sd, g, f"{agg}.gauss_pred.0", 2 * num_experts, D)
        _lin(sd, g, f"{agg}.router.0", num_experts, D)
        for e in range(num_experts):
            _lin(sd, g, f"{agg}.experts.{e}.0", D // 2, D)
            _lin(sd, g, f"{agg}.experts.{e}.2", D, D // 2)
    _lin(sd, g, "head", 42, D)
    if text_tower:
        W = text_width
        q = "quest_encoder"
        sd[q + ".positional_embedding"] = torch.randn(77, W, generator=g) * 0.01
        sd[q + ".text_projection"] = torch.randn(W, W, generator=g) * W ** -0.5
        sd[q + ".logit_scale"] = torch.tensor(2.6593)
        sd[q + ".token_embedding.weight"] = torch.randn(49408, W, generator=g) * 0.02
        _ln(sd, g, q + ".ln_final", W)
        for l in range(text_layers):
            r = f"{q}.transformer.resblocks.{l}"
            sd[r + ".attn.in_proj_weight"] = torch.randn(3 * W, W, generator=g) * W ** -0.5
            sd[r + ".attn.in_proj_bias"] = torch.randn(3 * W, generator=g) * 0.02
            sd[r + ".attn.out_proj.weight"] = torch.randn(W, W, generator=g) * (W ** -0.5) * (2 * text_layers) ** -0.5
            sd[r + ".attn.out_proj.bias"] = torch.randn(W, generator=g) * 0.02
            _ln(sd, g, r + ".ln_1", W)
            sd[r + ".mlp.c_fc.weight"] = torch.randn(4 * W, W, generator=g) * (2 * W) ** -0.5
            sd[r + ".mlp.c_fc.bias"] = torch.randn(4 * W, generator=g) * 0.02
            sd[r + ".mlp.c_proj.weight"] = torch.randn(W, 4 * W, generator=g) * (W ** -0.5) * (2 * text_layers) ** -0.5
            sd[r + ".mlp.c_proj.bias"] = torch.randn(W, generator=g) * 0.02
            _ln(sd, g, r + ".ln_2", W)
    return sd


def make_batch(seed: int, B: int, T: int = 60, P: int = 14, audio_dim: int = 128, video_dim: int = 768,
               patch_dim: int = 1024, with_patch: bool = True) -> Dict[str, torch.Tensor]:
    """tokens [SOT, L ids in [1,49405], EOT, 0...], L~U[5,20); features N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    quest = torch.zeros(B, 77, dtype=torch.int32)
    lens = torch.randint(5, 20, (B,), generator=g)
    for b in range(B):
        L = int(lens[b])
        quest[b, 0] = 49406
        quest[b, 1:1 + L] = torch.randint(1, 49406, (L,), generator=g, dtype=torch.int32)
        quest[b, 1 + L] = 49407
    out = {
        "quest": quest,
        "audio": torch.randn(B, T, audio_dim, generator=g),
        "video": torch.randn(B, T, video_dim, generator=g),
    }
    if with_patch:
        out["patch"] = torch.randn(B, T, P, patch_dim, generator=g)
    out["label"] = torch.randint(0, 42, (B,), generator=g)
    out["qtype_label"] = torch.randint(0, 9, (B,), generator=g)
    return out


def make_vit_blocks(seed: int = 717, n_blocks: int = 2, dim: int = 1024, hidden: int = 4096) -> Dict[str, torch.Tensor]:
    """Seeded weights of `n_blocks` timm ViT blocks (names blocks.<i>.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*), for the
    ToMe-patched block of SURVEY 8f-N4 (ViT-L/16: dim 1024, 16 heads, mlp 4096)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for i in range(n_blocks):
        p = f"blocks.{i}"
        _ln(sd, g, p + ".norm1", dim)
        _lin(sd, g, p + ".attn.qkv", 3 * dim, dim, gain=1.0)
        _lin(sd, g, p + ".attn.proj", dim, dim, gain=1.0)
        _ln(sd, g, p + ".norm2", dim)
        _lin(sd, g, p + ".mlp.fc1", hidden, dim)
        _lin(sd, g, p + ".mlp.fc2", dim, hidden, gain=1.0)
    return sd
