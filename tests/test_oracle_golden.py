"""CPU: the oracle restatement vs fixtures produced by the REAL reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import qat_oracle as O
from oracle import synth, tome_oracle as TO
from oracle.make_golden import FWD_CASES, sub

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def relerr(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", ["fwd_b8", "fwd_b3", "fwd_b1", "fwd_k5", "fwd_stress"])
def test_forward_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T, E, K, wseed, bseed = FWD_CASES[name]
    sd = synth.make_state_dict(wseed, num_experts=E)
    batch = synth.make_batch(bseed, B, T)
    with torch.no_grad():
        out, aux = O.qa_tiger_forward(sd, batch, K=K, E=E, with_aux=True)
    assert relerr(out, g["out"]) < 2e-5
    assert out.argmax(1).tolist() == g["argmax"].tolist()
    for agg, key in (("at_aggregator", "at"), ("vt_aggregator", "vt")):
        assert aux[key]["topk_inds"].tolist() == g[agg + ".topk_inds"].astype(np.int64).tolist()
        assert relerr(aux[key]["topk_probs"], g[agg + ".topk_probs"]) < 1e-5
        assert relerr(aux[key]["gauss_weight"], g[agg + ".gauss_weight"]) < 1e-5
    for k in ("quest768", "words768", "audio_x", "video_x", "a_patch", "v_patch"):
        assert relerr(sub(aux[k]), g[k]) < 2e-5, k
    for k in ("a_global", "ap_global", "vp_global", "fusion1", "fusion2"):
        assert relerr(aux[k], g[k]) < 2e-5, k


@pytest.mark.parametrize("name", ["fwd_t150"])
def test_forward_matches_reference_logits_only(name):
    """Goldens that hold logits + router ranks only (B > 8): here the frame count past one 128-token attention tile."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T, E, K, wseed, bseed = FWD_CASES[name]
    sd = synth.make_state_dict(wseed, num_experts=E)
    batch = synth.make_batch(bseed, B, T)
    with torch.no_grad():
        out, aux = O.qa_tiger_forward(sd, batch, K=K, E=E, with_aux=True)
    assert relerr(out, g["out"]) < 2e-5
    assert out.argmax(1).tolist() == g["argmax"].tolist()
    for agg, key in (("at_aggregator", "at"), ("vt_aggregator", "vt")):
        assert aux[key]["topk_inds"].tolist() == g[agg + ".topk_inds"].astype(np.int64).tolist()


def test_batch_coupling_quirk_is_reproduced():
    """sample 0 alone differs from sample 0 inside a batch."""
    g = np.load(os.path.join(GOLDEN, "fwd_b3.npz"))
    sd = synth.make_state_dict(713, num_experts=7)
    batch = synth.make_batch(2, 3, 60)
    one = {k: v[:1] for k, v in batch.items()}
    with torch.no_grad():
        o1 = O.qa_tiger_forward(sd, one)
    assert relerr(o1, g["out"][:1]) > 1e-3


def test_blocks_match_reference():
    g = np.load(os.path.join(GOLDEN, "blocks.npz"))
    sd = synth.make_state_dict(716, num_experts=7, text_tower=False)
    gen = torch.Generator().manual_seed(6)
    B, T, D = 4, 60, 512
    audio, video = torch.randn(B, T, D, generator=gen), torch.randn(B, T, D, generator=gen)
    words = torch.randn(B, 77, D, generator=gen)
    a, v, (aw, vw) = O.avq_cross_attn(sd, "crs_attn", audio, video, words, visualize=True)
    for got, key in ((a, "avq_a"), (v, "avq_v"), (aw, "avq_aw"), (vw, "avq_vw")):
        assert relerr(sub(got), g[key]) < 1e-5, key
    for tag, (B2, T2, E2, K2) in {"g7": (6, 60, 7, 7), "g10": (5, 60, 10, 5), "g14": (4, 120, 14, 14),
                                  "g1": (1, 60, 7, 7)}.items():
        x = torch.randn(T2, B2, E2, D, generator=gen)
        rl = torch.randn(B2, E2, generator=gen)
        gr = torch.randn(B2, 2 * E2, generator=gen)
        out, aux = O.gauss_reduce(x, rl, gr, K2)
        assert aux["topk_inds"].tolist() == g[tag + ".topk_inds"].astype(np.int64).tolist()
        assert relerr(aux["topk_probs"], g[tag + ".topk_probs"]) < 1e-6
        assert relerr(aux["gauss_weight"], g[tag + ".gauss_weight"]) < 1e-6
        assert relerr(out, g[tag + ".out"]) < 1e-5
        # the coef refactor is the same function
        coef = O.moe_coef(aux["gauss_weight"], aux["topk_inds"], aux["topk_probs"], E2)
        out2 = torch.einsum("bte,tbec->bc", coef, x)
        assert relerr(out2, g[tag + ".out"]) < 1e-5


def test_tome_chain_matches_reference():
    g = np.load(os.path.join(GOLDEN, "tome.npz"))
    gen = torch.Generator().manual_seed(42)
    x = torch.randn(2, 577, 1024, generator=gen)
    xo, size, group, steps = TO.chain(x)
    assert len(steps) == int(g["chain.steps"]) == 23
    for k, (unm, src, dst) in enumerate(steps):
        assert unm.tolist() == g[f"chain.{k}.unm"].astype(np.int64).tolist(), k
        assert src.tolist() == g[f"chain.{k}.src"].astype(np.int64).tolist(), k
        assert dst.tolist() == g[f"chain.{k}.dst"].astype(np.int64).tolist(), k
    assert xo.shape == (2, 14, 1024)
    assert relerr(xo, g["chain.x"]) < 1e-5
    assert torch.equal(size, torch.as_tensor(g["chain.size"]))
    assert float(size.sum()) == 2 * 577
    assert group.tolist() == g["chain.group"].astype(np.int64).tolist()


@pytest.mark.parametrize("nm", ["s_a", "s_b", "s_c", "s_d", "s_e", "s_f"])
def test_tome_single_steps(nm):
    g = np.load(os.path.join(GOLDEN, "tome.npz"))
    n, t, c, r, cls = g[nm + ".cfg"].tolist()
    metric, x = torch.as_tensor(g[nm + ".metric"]), torch.as_tensor(g[nm + ".x"])
    res = TO.match(metric, r, class_token=bool(cls))
    assert res is not None
    for got, key in zip(res[:3], ("unm", "src", "dst")):
        assert got.tolist() == g[f"{nm}.{key}"].astype(np.int64).tolist(), key
    for fast in (True, False):
        xo, so = TO.merge_wavg(x, None, res[:3], fast=fast)
        assert relerr(xo, g[nm + ".xo"]) < 1e-6
        assert torch.equal(so, torch.as_tensor(g[nm + ".so"]))


@pytest.mark.parametrize("nm", ["d_a", "d_b", "d_c", "d_d", "d_e"])
def test_tome_single_steps_with_distill_token(nm):
    """merge.py:56-57, 77-78: the distillation token (b-token 0) is never a merge destination and the merged sequence starts
    [unm_0, dst_0, ...]; goldens from the real bipartite_soft_matching(distill_token=True)."""
    g = np.load(os.path.join(GOLDEN, "tome_distill.npz"))
    n, t, c, r, cls = g[nm + ".cfg"].tolist()
    metric, x = torch.as_tensor(g[nm + ".metric"]), torch.as_tensor(g[nm + ".x"])
    res = TO.match(metric, r, class_token=bool(cls), distill_token=True)
    assert res is not None
    for got, key in zip(res[:3], ("unm", "src", "dst")):
        assert got.tolist() == g[f"{nm}.{key}"].astype(np.int64).tolist(), key
    assert not (res[2] == 0).any()                                   # nothing merges into the distillation token
    for fast in (True, False):
        xo, so = TO.merge_wavg(x, None, res[:3], fast=fast, distill_token=True)
        assert relerr(xo, g[nm + ".xo"]) < 1e-6
        assert torch.equal(so, torch.as_tensor(g[nm + ".so"]))
    group = TO.group_ids_update(torch.arange(t)[None, :].expand(n, t).contiguous(), t, res[:3], distill_token=True)
    assert group.tolist() == g[nm + ".group"].astype(np.int64).tolist()


@pytest.mark.parametrize("nm", ["b197", "b577"])
def test_tome_vit_block_matches_reference(nm):
    """the oracle's ToMeAttention / ToMeBlock restatement against the real tome/patch/timm.py (goldens minted with
    timm stubbed, oracle/make_golden.py `tome_block`): attention output, metric = k.mean(heads), sizes, group ids, block output."""
    g = np.load(os.path.join(GOLDEN, "tome_block.npz"))
    N, t, nb = (int(v) for v in g[f"{nm}.cfg"][:3])
    rs = [int(v) for v in g[f"{nm}.cfg"][3:]]
    sd = synth.make_vit_blocks(717, 2)
    x = torch.randn(N, t, 1024, generator=torch.Generator().manual_seed(1000 + t))
    size, group = None, torch.arange(t)[None, :].expand(N, t).contiguous()
    with torch.no_grad():
        for i in range(nb):
            p = f"blocks.{i}"
            xa, metric = TO.tome_attention(sd, p + ".attn", TO._ln(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6), size, 16)
            assert relerr(sub(xa), g[f"{nm}.{i}.attn"]) < 2e-5
            assert relerr(metric, g[f"{nm}.{i}.metric"]) < 1e-5
            x, size, metric2, idx, group = TO.tome_block(sd, p, x, size, rs[i], 16, group=group)
            assert relerr(sub(x, 8192), g[f"{nm}.{i}.x"]) < 2e-5
            assert torch.equal(size, torch.as_tensor(g[f"{nm}.{i}.size"]))
    assert relerr(sub(x, 65536), g[f"{nm}.x_final"]) < 2e-5
    assert group.tolist() == g[f"{nm}.group"].astype(np.int64).tolist()
