"""TEST INFRASTRUCTURE: a minimal stand-in for `timm.models.vision_transformer` (timm is not installed in this image) so that
the REAL reference file /root/reference/src/tome/patch/timm.py can be imported and its ToMeAttention / ToMeBlock forwards run on
CPU to mint golden vectors (oracle/make_golden.py `tome_block`).  Only the attributes those forwards touch are provided, with
timm's own names and semantics (Attention: qkv / scale / attn_drop / proj / proj_drop / num_heads; Block: norm1 / attn /
drop_path / norm2 / mlp with fc1 -> GELU -> fc2).  Nothing here is on the product path."""
import sys
import types

import torch.nn as nn


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads, qkv_bias)
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class VisionTransformer(nn.Module):
    pass


def install():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.Attention, vt.Block, vt.VisionTransformer = Attention, Block, VisionTransformer
    timm.models, models.vision_transformer = models, vt
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", models)
    sys.modules.setdefault("timm.models.vision_transformer", vt)
