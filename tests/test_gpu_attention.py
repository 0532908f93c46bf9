"""GPU parity of the attention cores (qat_attention_fwd) against an fp64 torch restatement of the MHA slow path
(q*hd^-0.5, softmax over all keys, no masks except the text tower's causal one), over every
(Lq, Lk) class the hot path uses and the padding / tail cases around them."""
import pytest
import torch

from gpu_util import dev, relerr

pytestmark = pytest.mark.gpu


def ref_attention(q, k, v, causal=False):
    q, k, v = (t.double().permute(0, 2, 1, 3) for t in (q, k, v))          # [N,H,L,64]
    s = (q * 0.125) @ k.transpose(-1, -2)
    if causal:
        s = s + torch.full(s.shape[-2:], float("-inf"), device=s.device, dtype=s.dtype).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[2], -1)
    return o, p.mean(dim=1)


# (N, H, Lq, Lk): AVQ 60x60 / 60x77 / stress 120x120, text tower 77x77 (12 heads), TempMoE 1x60, grounding 1x2,
# patch 14x14 and 2x14, plus padding / tail shapes of every core
F32_SHAPES = [(3, 8, 60, 60), (2, 8, 60, 77), (2, 8, 120, 120), (2, 12, 77, 77), (5, 8, 1, 60), (4, 8, 1, 2), (6, 8, 14, 14),
              (6, 8, 2, 14), (2, 8, 33, 17), (2, 8, 64, 64), (1, 8, 65, 128), (3, 4, 7, 5), (2, 8, 100, 81), (2, 8, 16, 40)]


@pytest.mark.parametrize("N,H,Lq,Lk", F32_SHAPES)
def test_attention_fp32_cores(N, H, Lq, Lk):
    from qa_tiger_b200 import _cabi
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + Lq * 10 + Lk)
    q = torch.randn(N, Lq, H, 64, device=dev(), generator=g)
    k = torch.randn(N, Lk, H, 64, device=dev(), generator=g)
    v = torch.randn(N, Lk, H, 64, device=dev(), generator=g)
    ro, rw = ref_attention(q, k, v)
    o = _cabi.attention(q, k, v)
    assert relerr(o, ro) < 2e-6
    if Lq > 2:                       # the <= 2-query streaming core has no weights output (visualize is an AVQ feature)
        o2, w = _cabi.attention(q, k, v, need_weights=True)
        assert relerr(o2, ro) < 2e-6 and relerr(w, rw) < 2e-6


@pytest.mark.parametrize("N,H,L", [(2, 12, 77), (3, 8, 60), (1, 8, 128), (2, 8, 33)])
def test_attention_causal(N, H, L):
    from qa_tiger_b200 import _cabi
    g = torch.Generator(device="cuda").manual_seed(L)
    q, k, v = (torch.randn(N, L, H, 64, device=dev(), generator=g) for _ in range(3))
    ro, _ = ref_attention(q, k, v, causal=True)
    assert relerr(_cabi.attention(q, k, v, causal=True), ro) < 2e-6


@pytest.mark.parametrize("N,H,Lq,Lk", [(50, 8, 14, 14), (50, 8, 2, 14), (7, 8, 16, 16), (9, 8, 5, 5), (4, 4, 1, 3), (3, 8, 14, 9), (2, 1, 14, 14)])
def test_attention_fp16_small_core(N, H, Lq, Lk):
    """fp16 planes in, fp16 out: exact products of the fp16 inputs, fp32 softmax, P rounded to fp16 once."""
    from qa_tiger_b200 import _cabi
    g = torch.Generator(device="cuda").manual_seed(N + Lq + Lk)
    q = torch.randn(N, Lq, H, 64, device=dev(), generator=g).half()
    kv = torch.randn(N, Lk, 2, H, 64, device=dev(), generator=g).half()
    ro, _ = ref_attention(q, kv[:, :, 0], kv[:, :, 1])
    o = _cabi.attention(q, kv, None)
    assert o.dtype == torch.float16 and o.shape == (N, Lq, H * 64)
    assert relerr(o, ro) < 6e-4
