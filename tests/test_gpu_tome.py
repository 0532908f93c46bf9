"""GPU ToMe: merge indices must be BIT-EXACT against the reference goldens."""
import numpy as np
import pytest
import torch

from oracle import tome_oracle as TO

from gpu_util import dev, golden, relerr

pytestmark = pytest.mark.gpu


def _tm():
    from qa_tiger_b200.tome import merge as tm
    return tm


@pytest.mark.parametrize("nm", ["s_a", "s_b", "s_c", "s_d", "s_e", "s_f"])
def test_single_steps_bit_exact(nm):
    tm = _tm()
    g = golden("tome")
    n, t, c, r, cls = g[nm + ".cfg"].tolist()
    metric, x = torch.as_tensor(g[nm + ".metric"]).to(dev()), torch.as_tensor(g[nm + ".x"]).to(dev())
    merge, unmerge = tm.bipartite_soft_matching(metric, r, class_token=bool(cls))
    for key in ("unm", "src", "dst"):
        got = getattr(merge, key + "_idx")
        assert got.dtype == torch.int64 and got.shape[-1] == 1
        assert got[..., 0].cpu().tolist() == g[f"{nm}.{key}"].astype(np.int64).tolist(), key
    xo, so = tm.merge_wavg(merge, x, None)
    assert torch.equal(xo.cpu(), torch.as_tensor(g[nm + ".xo"]))          # bit-exact values too
    assert torch.equal(so.cpu(), torch.as_tensor(g[nm + ".so"]))
    # bare closure modes vs the oracle's explicit scatter
    idx = [i[..., 0].cpu() for i in (merge.unm_idx, merge.src_idx, merge.dst_idx)]
    for mode in ("sum", "amax"):
        assert torch.equal(merge(x, mode=mode).cpu(), TO.merge_apply(x.cpu(), *idx, mode=mode))
    # unmerge(merge(x)) keeps every destination/unmerged token where it was
    back = unmerge(merge(x, mode="sum"))
    assert back.shape == x.shape


def test_chain_577_to_14_bit_exact():
    tm = _tm()
    g = golden("tome")
    x = torch.randn(2, 577, 1024, generator=torch.Generator().manual_seed(42)).to(dev())
    xo, size, group, steps = tm.chain(x, trace=True)
    assert len(steps) == int(g["chain.steps"]) == 23
    for k, (unm, src, dst) in enumerate(steps):
        assert unm.cpu().tolist() == g[f"chain.{k}.unm"].astype(np.int64).tolist(), k
        assert src.cpu().tolist() == g[f"chain.{k}.src"].astype(np.int64).tolist(), k
        assert dst.cpu().tolist() == g[f"chain.{k}.dst"].astype(np.int64).tolist(), k
    assert xo.shape == (2, 14, 1024)
    assert torch.equal(xo.cpu(), torch.as_tensor(g["chain.x"]))
    assert torch.equal(size.cpu(), torch.as_tensor(g["chain.size"]))
    assert group.cpu().tolist() == g["chain.group"].astype(np.int64).tolist()


def test_merge_source_adjacency_matches_group_ids():
    tm = _tm()
    gen = torch.Generator().manual_seed(3)
    metric = torch.randn(2, 27, 64, generator=gen).to(dev())
    x = torch.randn(2, 27, 32, generator=gen).to(dev())
    merge, _ = tm.bipartite_soft_matching(metric, 25, class_token=True)
    source = tm.merge_source(merge, x, None)                              # [2,14,27] one-hot adjacency
    assert source.shape == (2, 14, 27) and float(source.sum()) == 2 * 27
    group = torch.arange(27, dtype=torch.int32, device=dev())[None].repeat(2, 1).contiguous()
    tm.merge_with(x, merge.idx, "wavg", None, group)
    assert source.argmax(dim=1).cpu().tolist() == group.cpu().long().tolist()


def test_chain_properties_many_frames():
    """Size-independent checks at a larger frame count: sizes sum to t0, cls stays first, idempotent & deterministic."""
    tm = _tm()
    N = 96
    x = torch.randn(N, 577, 1024, device=dev(), generator=torch.Generator(device="cuda").manual_seed(7))
    xo, size, group, _ = tm.chain(x)
    assert xo.shape == (N, 14, 1024)
    assert torch.equal(size.sum(dim=(1, 2)).cpu(), torch.full((N,), 577.0))
    assert int(group[:, 0].max()) == 0                                     # class token never merges
    counts = torch.zeros(N, 14, device=dev()).scatter_add_(1, group.long(), torch.ones(N, 577, device=dev()))
    assert torch.equal(counts, size[..., 0])
    xo2, size2, group2, _ = tm.chain(x)
    assert torch.equal(xo, xo2) and torch.equal(group, group2)
    # against the CPU oracle on a slice of frames (indices bit-exact, values bit-exact)
    ro, rs, rg, _ = TO.chain(x[:3].cpu())
    assert torch.equal(xo[:3].cpu(), ro) and rg.tolist() == group[:3].cpu().long().tolist()


def test_r_zero_and_errors():
    tm = _tm()
    m = torch.randn(1, 1, 64, device=dev())
    merge, unmerge = tm.bipartite_soft_matching(m, 25, class_token=True)
    assert merge is tm.do_nothing
    merge, _ = tm.bipartite_soft_matching(torch.randn(1, 3, 64, device=dev()), 2, class_token=True, distill_token=True)
    assert merge is tm.do_nothing                                  # (3 - 2 protected) // 2 = 0
    with pytest.raises(RuntimeError, match="C="):
        tm.bipartite_soft_matching(torch.randn(1, 8, 48, device=dev()), 2)


def test_chain_frames_side_streams_match_serial_chunks():
    """chain_frames over several chunks on two side streams (match kernels of one chunk overlapping merge kernels of another) returns
    exactly what the serial chunk loop returns: the chunks are independent."""
    from qa_tiger_b200.tome import merge as tm
    g = torch.Generator().manual_seed(5)
    x = torch.randn(23, 577, 1024, generator=g).cuda()
    a = tm.chain_frames(x, 23, chunk=5, want_group=True, streams=1)
    b = tm.chain_frames(x, 23, chunk=5, want_group=True, streams=2)
    c = tm.chain_frames(lambda lo, hi: x[lo:hi].clone(), 23, chunk=5, want_group=True, streams=3)
    torch.cuda.synchronize()
    for u, v, w in zip(a, b, c):
        assert torch.equal(u, v) and torch.equal(u, w)


@pytest.mark.parametrize("nm", ["d_a", "d_b", "d_c", "d_d", "d_e"])
def test_single_steps_with_distill_token_bit_exact(nm):
    """bipartite_soft_matching(distill_token=True) (merge.py:56-57, 77-78) against goldens of the real reference: indices, merge_wavg
    values and sizes, the bare mean-mode closure and the merge_source groups, all bit-exact."""
    tm = _tm()
    g = golden("tome_distill")
    n, t, c, r, cls = g[nm + ".cfg"].tolist()
    metric, x = torch.as_tensor(g[nm + ".metric"]).to(dev()), torch.as_tensor(g[nm + ".x"]).to(dev())
    merge, _ = tm.bipartite_soft_matching(metric, r, class_token=bool(cls), distill_token=True)
    for key in ("unm", "src", "dst"):
        assert getattr(merge, key + "_idx")[..., 0].cpu().tolist() == g[f"{nm}.{key}"].astype(np.int64).tolist(), key
    xo, so = tm.merge_wavg(merge, x, None)
    assert torch.equal(xo.cpu(), torch.as_tensor(g[nm + ".xo"]))
    assert torch.equal(so.cpu(), torch.as_tensor(g[nm + ".so"]))
    assert torch.equal(merge(x, mode="mean").cpu(), torch.as_tensor(g[nm + ".mean"]))
    group = torch.arange(t, dtype=torch.int32, device=dev())[None, :].expand(n, t).contiguous()
    tm.merge_with(x, merge.idx, "wavg", None, group=group, distill_token=True)
    assert group.cpu().tolist() == g[nm + ".group"].astype(np.int64).tolist()
