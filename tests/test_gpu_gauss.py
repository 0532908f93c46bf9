"""GPU parity of the fused Gaussian-expert kernel (`qat_gauss_reduce_fwd`) against
reference-generated goldens, the oracle, and size-independent properties at full size."""
import numpy as np
import pytest
import torch

from oracle import qat_oracle as O

from gpu_util import dev, golden, relerr

pytestmark = pytest.mark.gpu


def _gr():
    from qa_tiger_b200.models.modules import gauss_reduce
    return gauss_reduce


@pytest.mark.parametrize("tag", ["g7", "g10", "g14", "g1"])
@pytest.mark.parametrize("t_split", [0, 1, 2, 4, 8])
def test_golden_from_reference(tag, t_split):
    g = golden("blocks")
    B, T, E, K = g[tag + ".shape"].tolist()
    gen = torch.Generator().manual_seed(6)
    # regenerate the inputs exactly as oracle/make_golden.py:gen_blocks does
    _ = [torch.randn(4, 60, 512, generator=gen), torch.randn(4, 60, 512, generator=gen), torch.randn(4, 77, 512, generator=gen)]
    for tg, (B2, T2, E2, K2) in {"g7": (6, 60, 7, 7), "g10": (5, 60, 10, 5), "g14": (4, 120, 14, 14), "g1": (1, 60, 7, 7)}.items():
        x = torch.randn(T2, B2, E2, 512, generator=gen)
        rl = torch.randn(B2, E2, generator=gen)
        gr = torch.randn(B2, 2 * E2, generator=gen)
        if tg == tag:
            break
    out, (inds, probs, gw) = _gr()(x.to(dev()), rl.to(dev()), gr.to(dev()), K, t_split=t_split, return_aux=True)
    assert inds.cpu().tolist() == g[tag + ".topk_inds"].astype(np.int64).tolist()
    assert relerr(probs, g[tag + ".topk_probs"]) < 1e-6
    assert relerr(gw, g[tag + ".gauss_weight"]) < 2e-6
    assert relerr(out, g[tag + ".out"]) < 2e-6


@pytest.mark.parametrize("B,T,E,K,C", [(256, 60, 7, 7, 512), (33, 60, 7, 4, 512), (128, 120, 14, 14, 512), (7, 13, 3, 2, 64),
                                       (300, 60, 7, 7, 256), (2, 2, 1, 1, 512), (16, 60, 32, 9, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vs_oracle(B, T, E, K, C, dtype):
    gen = torch.Generator().manual_seed(B + T + E)
    x = torch.randn(T, B, E, C, generator=gen).to(dtype)
    rl, gr = torch.randn(B, E, generator=gen) * 2, torch.randn(B, 2 * E, generator=gen) * 2
    ref, aux = O.gauss_reduce(x.float(), rl, gr, K)
    out, (inds, probs, gw) = _gr()(x.to(dev()), rl.to(dev()), gr.to(dev()), K, return_aux=True)
    assert inds.cpu().tolist() == aux["topk_inds"].tolist()
    assert relerr(gw, aux["gauss_weight"]) < 2e-6
    assert relerr(out, ref) < 5e-6


def test_properties_full_size():
    """Config 2/5 sizes: linearity in X and the constant-input identity out[b,c] = sum_t sum_j coef[b,t,j]."""
    gr_fn = _gr()
    for (B, T, E, K) in [(256, 60, 7, 7), (128, 120, 14, 14)]:
        gen = torch.Generator(device="cuda").manual_seed(5)
        x = torch.randn(T, B, E, 512, device=dev(), generator=gen)
        y = torch.randn(T, B, E, 512, device=dev(), generator=gen)
        rl, gr = torch.randn(B, E, device=dev(), generator=gen), torch.randn(B, 2 * E, device=dev(), generator=gen)
        ox, oy = gr_fn(x, rl, gr, K), gr_fn(y, rl, gr, K)
        oxy = gr_fn(2 * x - 3 * y, rl, gr, K)
        assert relerr(oxy, 2 * ox - 3 * oy) < 1e-5
        ones, (inds, probs, gw) = gr_fn(torch.ones_like(x), rl, gr, K, return_aux=True)
        total = (gw * probs[:, :, None]).sum(dim=(1, 2))                      # K == E: every expert selected
        assert relerr(ones, total[:, None].expand(B, 512)) < 1e-5
        assert torch.equal(gr_fn(x, rl, gr, K), ox)                            # bit-reproducible
        assert float(gw.max()) == 1.0 and float(gw.amax(dim=2).min()) == 1.0   # peak normalisation


def test_bad_arguments_raise():
    x = torch.randn(4, 2, 3, 512, device=dev())
    with pytest.raises(RuntimeError, match="K"):
        _gr()(x, torch.randn(2, 3, device=dev()), torch.randn(2, 6, device=dev()), 5)
    with pytest.raises(RuntimeError):
        _gr()(x.cpu(), torch.randn(2, 3), torch.randn(2, 6), 2)
