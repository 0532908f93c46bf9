"""The library's A/B options (include/qat.h `qat_set_option`) select alternative kernels for the same
arithmetic -- the unfused forms that also serve the shapes the fused kernels do not cover.  These tests check each against
the default path on the same inputs."""
import pytest
import torch

from oracle import synth

from gpu_util import dev, relerr, sub_sd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return synth.make_state_dict(716, num_experts=7, text_tower=False)


def _patch_block(sd):
    from qa_tiger_b200.models import modules as M
    m = M.PatchSelecter(512, 8)
    m.load_state_dict(sub_sd(sd, "patch_selecter"))
    m = m.to(dev()).eval()
    m.precision = "bf16"
    proj = M.Projection(1024, 512)
    proj.load_state_dict(sub_sd(sd, "patch_proj"))
    proj = proj.to(dev()).eval()
    proj.precision = "bf16"
    return m, proj


@pytest.mark.parametrize("switch", ["patch_no_fused_cross", "patch_no_fused_attn"])
def test_patch_selecter_kernel_alternatives(sd, switch):
    """Unfused cross-attention, unfused self-attention: same fp16-plane arithmetic."""
    from qa_tiger_b200 import _cabi
    m, proj = _patch_block(sd)
    g = torch.Generator().manual_seed(7)
    raw = torch.randn(12, 60, 14, 1024, generator=g).to(dev())
    a, v = torch.randn(12, 60, 512, generator=g).to(dev()), torch.randn(12, 60, 512, generator=g).to(dev())
    with torch.no_grad():
        da, dv = m(raw, a, v, proj=proj)
        with _cabi.option(switch, 1):
            sa, sv = m(raw, a, v, proj=proj)
    assert _cabi.get_option(switch) == 0
    # identical operands and fp32 accumulation; only the softmax / store schedule differs (fp16 rounding of P and O)
    assert relerr(sa, da) < 5e-4 and relerr(sv, dv) < 5e-4


@pytest.mark.parametrize("switch,val", [("attn_fp32_in", 1), ("gemm_cg", 1)])
def test_avq_kernel_alternatives(sd, switch, val):
    """fp32 instead of plane-pair in-projections into the attention cores; single-CTA instead of CTA-pair GEMM tiles."""
    from qa_tiger_b200 import _cabi
    from qa_tiger_b200.models import modules as M
    m = M.AVQCrossAttn(512, 8)
    m.load_state_dict(sub_sd(sd, "crs_attn"))
    m = m.to(dev()).eval()
    m.precision = "bf16"
    g = torch.Generator().manual_seed(11)
    a, v, w = (torch.randn(20, 60, 512, generator=g).to(dev()), torch.randn(20, 60, 512, generator=g).to(dev()),
               torch.randn(20, 77, 512, generator=g).to(dev()))
    with torch.no_grad():
        da, dv = m(a, v, w)
        with _cabi.option(switch, val):
            sa, sv = m(a, v, w)
    assert relerr(sa, da) < 2e-5 and relerr(sv, dv) < 2e-5


def test_tome_match_kernels_agree():
    """Register-tiled and row-per-thread match kernels: the same indices and merged tokens, bit for bit."""
    from qa_tiger_b200 import _cabi
    from qa_tiger_b200.tome import merge as tm
    x = torch.randn(5, 577, 1024, generator=torch.Generator().manual_seed(3)).to(dev())
    xo, size, group, steps = tm.chain(x, trace=True)
    with _cabi.option("tome_match_rows", 1):
        xo2, size2, group2, steps2 = tm.chain(x, trace=True)
    assert torch.equal(xo, xo2) and torch.equal(size, size2) and torch.equal(group, group2)
    for s1, s2 in zip(steps, steps2):
        for t1, t2 in zip(s1, s2):
            assert torch.equal(t1, t2)


@pytest.mark.parametrize("B,T,E,K,two", [(40, 60, 7, 7, True), (40, 60, 7, 7, False), (12, 120, 14, 14, True), (37, 60, 10, 5, True)])
def test_moe_fused_reduce_matches_the_streaming_kernel(B, T, E, K, two):
    """TempMoE with M4 in the expert GEMM's epilogue (default) against first-layer GEMM + gauss_reduce_kernel launches: the same
    coefficients and products, another summation order (fp32)."""
    from qa_tiger_b200 import _cabi
    from qa_tiger_b200.models import modules as M
    sdm = synth.make_state_dict(720 + E, num_experts=E, text_tower=False)
    pfx = "vt_aggregator" if two else "at_aggregator"
    m = M.TempMoE(512, 8, topK=K, n_experts=E, vis_branch=two)
    m.load_state_dict(sub_sd(sdm, pfx))
    m = m.to(dev()).eval()
    m.precision = "bf16"
    m.keep_aux = True
    g = torch.Generator().manual_seed(B + T)
    q, d = torch.randn(B, 512, generator=g).to(dev()), torch.randn(B, T, 512, generator=g).to(dev())
    sub = [torch.randn(B, T, 512, generator=g).to(dev()), torch.randn(B, T, 512, generator=g).to(dev())] if two else None
    with torch.no_grad(), _cabi.option("moe_w2_simt", 1):        # the same fp32 second layer behind both forms
        fused = m(q, d, sub)
        aux_f = [t.clone() for t in m.last_aux]
        with _cabi.option("moe_no_fused_reduce", 1):
            plain = m(q, d, sub)
        aux_p = m.last_aux
    for x, y in zip(aux_f, aux_p):
        assert torch.equal(x, y)
    for x, y in zip(fused if two else [fused], plain if two else [plain]):
        assert relerr(x, y) < 2e-6
    # second expert layer on the split-K bf16x3 tensor-core path (default) against the fp32 CUDA-core GEMM: 16-bit operands, after the router
    with torch.no_grad():
        tcw2 = m(q, d, sub)
    for x, y in zip(tcw2 if two else [tcw2], fused if two else [fused]):
        assert relerr(x, y) < 3e-5


def test_patch_selecter_a_resident_matches_shared_memory_operands(sd):
    """Fused in-projection + attention kernels with the A operand in tensor memory (default, gemm_attn_ts.cuh) against the form that
    feeds both operands from shared memory (option gemm_attn_no_ts = 1): the same fp16 products accumulated in the same k order ->
    identical results; shapes with a single unit per CTA, CTAs that start in the middle of a row tile, and a ragged last tile."""
    from qa_tiger_b200 import _cabi
    m, proj = _patch_block(sd)
    g = torch.Generator().manual_seed(13)
    for B, T in ((1, 9), (1, 1), (3, 31), (12, 60), (40, 60)):     # 1, 1, 11, 80, 267 row tiles of 9 frames (x 8 heads)
        raw = torch.randn(B, T, 14, 1024, generator=g).to(dev())
        a, v = torch.randn(B, T, 512, generator=g).to(dev()), torch.randn(B, T, 512, generator=g).to(dev())
        with torch.no_grad():
            da, dv = m(raw, a, v, proj=proj)
            with _cabi.option("gemm_attn_no_ts", 1):
                sa, sv = m(raw, a, v, proj=proj)
        assert torch.equal(sa, da) and torch.equal(sv, dv), (B, T, relerr(da, sa), relerr(dv, sv))
