"""GPU parity of the three GEMM engines (fp32 FFMA, tcgen05 tf32, tcgen05 bf16x3) against an fp64 torch
reference, over tile tails in M, N and K, with bias / ReLU / residual epilogues and bf16-plane outputs."""
import pytest
import torch

from gpu_util import dev, relerr

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (1, 8, 32), (130, 72, 96), (1000, 512, 512), (4096, 1536, 512), (333, 3584, 512), (257, 256, 1856),
          (15360, 1792, 512), (2048, 512, 1024), (640, 1024, 768), (300, 512, 128)]


@pytest.mark.parametrize("kind,tol", [(0, 2e-6), (1, 1.5e-3), (2, 2e-5)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_kinds(kind, tol, M, N, K):
    from qa_tiger_b200 import _cabi
    if kind == 0 and K % 16:
        pytest.skip("SIMT engine wants K % 16 == 0")
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev(), generator=g)
    w = torch.randn(N, K, device=dev(), generator=g) * K ** -0.5
    b = torch.randn(N, device=dev(), generator=g)
    res = torch.randn(M, N, device=dev(), generator=g)
    lin = _cabi.Lin(w, b, _cabi.PREC_BF16)
    ref = a.double() @ w.double().t() + b.double()
    got = _cabi.gemm(kind, a, lin)
    assert relerr(got, ref) < tol
    got2, (hi, lo) = _cabi.gemm(kind, a, lin, relu=True, res=res, out_planes=True)
    ref2 = torch.relu(ref) + res.double()
    assert relerr(got2, ref2) < tol
    assert relerr(hi.float() + lo.float(), got2) < 1e-4
    assert torch.equal(hi, got2.to(torch.bfloat16))


@pytest.fixture
def opt():
    """set A/B options of the library for one test, restore them afterwards"""
    from qa_tiger_b200 import _cabi
    prev = {}

    def _set(name, value):
        prev.setdefault(name, _cabi.get_option(name))
        _cabi.set_option(name, value)
    yield _set
    for k, v in prev.items():
        _cabi.set_option(k, v)


@pytest.mark.parametrize("bn", [128, 256])
def test_gemm_tile_widths(bn, opt):
    from qa_tiger_b200 import _cabi
    opt("gemm_bn", bn)
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(777, 512, device=dev(), generator=g)
    w = torch.randn(1024, 512, device=dev(), generator=g) * 512 ** -0.5
    lin = _cabi.Lin(w, None, _cabi.PREC_BF16)
    ref = a.double() @ w.double().t()
    assert relerr(_cabi.gemm(2, a, lin), ref) < 2e-5
    assert relerr(_cabi.gemm(1, a, lin), ref) < 1.5e-3


F16_SHAPES = [(128, 128, 64), (1, 8, 64), (130, 72, 128), (1000, 512, 512), (4096, 1536, 512), (14 * 50, 1024, 512), (257, 256, 1856),
              (57344, 512, 512), (126, 1536, 512)]


@pytest.mark.parametrize("M,N,K", F16_SHAPES)
def test_gemm_f16_kind(M, N, K):
    """tcgen05 kind::f16 on single fp16 planes: exact products of the fp16-rounded operands, fp32 accumulate."""
    from qa_tiger_b200 import _cabi
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev(), generator=g).half()
    w = torch.randn(N, K, device=dev(), generator=g) * K ** -0.5
    b = torch.randn(N, device=dev(), generator=g)
    res = torch.randn(M, N, device=dev(), generator=g).half()
    lin = _cabi.Lin(w, b, _cabi.PREC_BF16, f16=True)
    ref = a.double() @ lin.h.double().t() + b.double()
    got = _cabi.gemm(_cabi.GEMM_F16, a, lin)                       # generic epilogue, fp32 out
    assert relerr(got, ref) < 2e-6
    got_h = _cabi.gemm(_cabi.GEMM_F16, a, lin, out_f16=True)       # TMA-store epilogue
    assert got_h.dtype == torch.float16
    assert torch.equal(got_h, got.half())
    ref2 = torch.relu(ref) + res.double()
    got2 = _cabi.gemm(_cabi.GEMM_F16, a, lin, relu=True, res=res, out_f16=True)
    assert relerr(got2, ref2) < 4e-4
    got3 = _cabi.gemm(_cabi.GEMM_F16, a, lin, relu=True, res=res)  # fp16 residual through the generic epilogue
    assert relerr(got3, ref2) < 2e-6
    assert torch.equal(got2, got3.half())


@pytest.mark.parametrize("M,N,K", [(1000, 512, 1024), (14 * 300, 512, 1024), (77, 256, 96)])
def test_gemm_tf32_to_f16_plane(M, N, K):
    """patch_proj form: tf32 contraction of fp32 operands whose only output is an fp16 plane (TMA store)."""
    from qa_tiger_b200 import _cabi
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev(), generator=g)
    w = torch.randn(N, K, device=dev(), generator=g) * K ** -0.5
    b = torch.randn(N, device=dev(), generator=g)
    lin = _cabi.Lin(w, b, _cabi.PREC_BF16)
    full = _cabi.gemm(_cabi.GEMM_TF32, a, lin)
    got = _cabi.gemm(_cabi.GEMM_TF32, a, lin, out_f16=True)
    assert torch.equal(got, full.half())


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M", [129, 777, 1024, 40 * 128 + 5])
def test_gemm_cta_pairs(cg, M, opt):
    """cta_group::2 (256-row pair tiles, W rows split across the pair) against single-CTA tiles: every kind, odd and even
    numbers of 128-row tiles (a tail pair is half empty), bias / ReLU / residual epilogues, plane outputs."""
    from qa_tiger_b200 import _cabi
    opt("gemm_cg", cg)
    N, K = 768, 320
    g = torch.Generator(device="cuda").manual_seed(M)
    a = torch.randn(M, K, device=dev(), generator=g)
    w = torch.randn(N, K, device=dev(), generator=g) * K ** -0.5
    b = torch.randn(N, device=dev(), generator=g)
    res = torch.randn(M, N, device=dev(), generator=g)
    lin = _cabi.Lin(w, b, _cabi.PREC_BF16, f16=True)
    ref = a.double() @ w.double().t() + b.double()
    ref2 = torch.relu(ref) + res.double()
    for kind, tol in ((_cabi.GEMM_TF32, 1.5e-3), (2, 2e-5)):
        assert relerr(_cabi.gemm(kind, a, lin), ref) < tol
        got2, (hi, lo) = _cabi.gemm(kind, a, lin, relu=True, res=res, out_planes=True)
        assert relerr(got2, ref2) < tol
        assert torch.equal(hi, got2.to(torch.bfloat16))
    ah = a.half()
    refh = ah.double() @ lin.h.double().t() + b.double()
    got = _cabi.gemm(_cabi.GEMM_F16, ah, lin)
    assert relerr(got, refh) < 2e-6
    assert torch.equal(_cabi.gemm(_cabi.GEMM_F16, ah, lin, out_f16=True), got.half())
    rh = res.half()
    got3 = _cabi.gemm(_cabi.GEMM_F16, ah, lin, relu=True, res=rh)
    assert torch.equal(_cabi.gemm(_cabi.GEMM_F16, ah, lin, relu=True, res=rh, out_f16=True), got3.half())


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("S,T,E,rep", [(5, 60, 7, 1), (64, 60, 7, 2), (3, 120, 14, 1), (9, 32, 2, 1), (7, 100, 3, 1), (4, 128, 1, 1), (12, 50, 4, 2),
                                       (257, 60, 7, 1)])
def test_gemm_reduce_epilogue(cg, S, T, E, rep, opt):
    """Expert first-layer GEMM with M4 in its epilogue (gemm_tc.cuh EPI = 4): out[s, n] = sum_t coef[row, n // Ch] relu(a W^T + b).
    Samples of 32..128 rows (tiles of whole samples, warps straddling two samples, tail tiles / half-empty tail pairs),
    the coefficient table shared by `rep` row blocks (the two vt branches), single CTAs and CTA pairs."""
    from qa_tiger_b200 import _cabi
    opt("gemm_cg", cg)
    Ch, K = 256, 192
    M, N = S * T, E * Ch
    g = torch.Generator(device="cuda").manual_seed(S * 1000 + T)
    a = torch.randn(M, K, device=dev(), generator=g)
    w = torch.randn(N, K, device=dev(), generator=g) * K ** -0.5
    b = torch.randn(N, device=dev(), generator=g)
    assert S % rep == 0 or rep == 1
    R = M // rep if S % rep == 0 else M
    coef = torch.rand(R, E, device=dev(), generator=g)
    lin = _cabi.Lin(w, b, _cabi.PREC_BF16)
    got = _cabi.gemm_reduce(a, lin, coef, T, Ch)
    h = torch.relu(a.double() @ w.double().t() + b.double())                    # [M, N]
    cf = coef.double()[torch.arange(M, device=dev()) % R]                        # [M, E]
    ref = (h.view(S, T, E, Ch) * cf.view(S, T, E, 1)).sum(1).reshape(S, N)
    assert got.shape == (S, N)
    assert relerr(got, ref) < 2e-5
    # bit-reproducible: no atomics anywhere in the reduction
    assert torch.equal(got, _cabi.gemm_reduce(a, lin, coef, T, Ch))


@pytest.mark.parametrize("nf,cross", [(1, False), (9, False), (40, False), (1000, False), (9, True), (1000, True)])
def test_fused_projection_attention_matches_fp64(nf, cross):
    """qat_gemm_attn_fwd (PatchSelecter's fused in-projection + 14-token attention / K|V projection + two-query attention, A operand
    in tensor memory) against the same fp16 operands contracted in fp64: the projection is exact up to fp32 accumulation, then rounded
    to fp16 (as the kernel stores it in its shared-memory tile), softmax in fp32, P and O rounded to fp16."""
    from qa_tiger_b200 import _cabi
    P, H, D, K = 14, 8, 512, 512
    g = torch.Generator().manual_seed(nf + int(cross))
    x = (torch.randn(nf * P, K, generator=g) * 0.5).half()
    w = (torch.randn((2 if cross else 3) * D, K, generator=g) * 0.05).half()
    b = torch.randn((2 if cross else 3) * D, generator=g) * 0.1
    q = torch.randn(nf, 2, D, generator=g).half() if cross else None
    o = _cabi.gemm_attn(x.cuda(), w.cuda(), b.cuda(), P, H, q=None if q is None else q.cuda()).float().cpu()
    proj = (x.double() @ w.double().t() + b.double()).half().double()                       # the tile is stored as fp16
    if cross:
        qq, kk, vv = q.double().view(nf, 2, H, 64), proj[:, :D].view(nf, P, H, 64), proj[:, D:].view(nf, P, H, 64)
    else:
        qq, kk, vv = (proj[:, i * D:(i + 1) * D].view(nf, P, H, 64) for i in range(3))
    s_ = torch.einsum("nihd,njhd->nhij", qq, kk) * 0.125
    ref = torch.einsum("nhij,njhd->nihd", torch.softmax(s_, -1), vv).reshape(o.shape)
    assert relerr(o, ref) < 2e-3, relerr(o, ref)                                            # P and O are fp16 (11 bits)
