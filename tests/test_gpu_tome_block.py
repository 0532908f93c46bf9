"""GPU parity of the ToMe-patched ViT block (tome/patch/timm.py ToMeAttention / ToMeBlock) through
qat_tome_attention_fwd / qat_tome_block_fwd, and of the long-sequence attention core with a per-key bias behind it."""
import numpy as np
import pytest
import torch

from oracle import synth, tome_oracle as TO
from oracle.make_golden import sub

from gpu_util import dev, golden, relerr

pytestmark = pytest.mark.gpu


def _blocks(prec, n=2):
    from qa_tiger_b200.tome import patch as P
    sd = synth.make_vit_blocks(717, n)
    out = []
    for i in range(n):
        b = P.ToMeBlock(1024, 16)
        b.load_state_dict({k[len(f"blocks.{i}."):]: v for k, v in sd.items() if k.startswith(f"blocks.{i}.")})
        b = b.to(dev()).eval()
        b.precision = prec
        b.attn.precision = prec
        out.append(b)
    return sd, out


@pytest.mark.parametrize("Lq,Lk,H,bias", [(577, 577, 16, True), (130, 300, 8, False), (64, 129, 2, True), (27, 27, 16, True), (200, 64, 4, True)])
def test_long_attention_core(Lq, Lk, H, bias):
    """online-softmax core (attn_long.cuh): Lk > 128 and / or a per-key bias, against fp64 torch"""
    from qa_tiger_b200 import _cabi
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    N = 3
    q, k, v = (torch.randn(N, L, H, 64, generator=g).to(dev()) for L in (Lq, Lk, Lk))
    kb = (torch.rand(N, Lk, generator=g) * 3 + 1).log().to(dev()) if bias else None
    o = _cabi.attention(q, k, v, key_bias=kb)
    s = torch.einsum("nihd,njhd->nhij", q.double(), k.double()) * 0.125
    if bias:
        s = s + kb.double()[:, None, None, :]
    ref = torch.einsum("nhij,njhd->nihd", s.softmax(-1), v.double()).reshape(N, Lq, H * 64)
    assert relerr(o, ref) < 3e-6


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 1e-4)])
def test_tome_attention_matches_reference(prec, tol):
    """ToMeAttention.forward: proportional attention + metric = k.mean(heads) against the reference golden (block 1 of b197:
    sizes != 1 after the first merge) and the oracle"""
    g = golden("tome_block")
    sd, blocks = _blocks(prec)
    N, t = 2, 197
    x = torch.randn(N, t, 1024, generator=torch.Generator().manual_seed(1000 + t))
    with torch.no_grad():
        x1, size1, _, _, _ = TO.tome_block(sd, "blocks.0", x, None, 25, 16)
        p = "blocks.1"
        y = TO._ln(x1, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6)
        ra, rm = TO.tome_attention(sd, p + ".attn", y, size1, 16)
        ga, gm = blocks[1].attn(y.to(dev()).contiguous(), size1.to(dev()))
    assert relerr(ga, ra) < tol and relerr(gm, rm) < tol
    assert relerr(sub(ga.cpu()), g["b197.1.attn"]) < 2 * tol


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 2e-4)])
@pytest.mark.parametrize("nm", ["b197", "b577"])
def test_tome_block_stack_matches_reference(nm, prec, tol):
    """ToMeBlock x n through the drop-in stack (r schedule, size and source bookkeeping): block outputs and sizes against the
    reference goldens; merge indices bit-exact against the oracle's matching run on the SAME metric the CUDA block produced."""
    from qa_tiger_b200.tome import patch as P
    g = golden("tome_block")
    N, t, nb = (int(v) for v in g[f"{nm}.cfg"][:3])
    rs = [int(v) for v in g[f"{nm}.cfg"][3:]]
    sd, blocks = _blocks(prec, nb)
    for b in blocks:
        b.keep_indices = True
    stack = P.ToMeStack(blocks, r=rs, trace_source=True)
    x = torch.randn(N, t, 1024, generator=torch.Generator().manual_seed(1000 + t)).to(dev())
    with torch.no_grad():
        y = stack(x)
    assert relerr(sub(y.cpu(), 65536), g[f"{nm}.x_final"]) < tol
    assert torch.equal(stack._tome_info["size"].cpu(), torch.as_tensor(g[f"{nm}.{nb - 1}.size"])) or prec == "bf16"
    for i, b in enumerate(blocks):
        assert relerr(b.last_metric, g[f"{nm}.{i}.metric"]) < tol
        res = TO.match(b.last_metric.cpu(), rs[i], class_token=True)
        for got, want in zip(b.last_indices, res[:3]):
            assert torch.equal(got.cpu().long(), want.reshape(got.shape).long())
    if prec == "fp32":
        # fp32 mode reproduces the reference's indices themselves on these seeds, hence its source groups
        assert stack._tome_info["source"].cpu().tolist() == g[f"{nm}.group"].astype(np.int64).tolist()
        assert torch.equal(P.source_matrix(stack._tome_info["source"], y.shape[1]).argmax(1).cpu(), stack._tome_info["source"].cpu().long())


def test_tome_block_without_merge_or_sizes():
    """r = 0 and size = None: a plain ViT block (no bias, no merge), size stays None as in the reference"""
    from qa_tiger_b200.tome import patch as P
    sd, blocks = _blocks("fp32", 1)
    stack = P.ToMeStack(blocks, r=0)
    x = torch.randn(2, 50, 1024, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref, size, *_ = TO.tome_block(sd, "blocks.0", x, None, 0, 16)
        y = stack(x.to(dev()))
    assert y.shape == (2, 50, 1024) and stack._tome_info["size"] is None and size is None
    assert relerr(y, ref) < 2e-5
