"""CPU, world_size 2 over gloo: the data-parallel host logic (contiguous sharding, packed counters, the single
all_gather + all_reduce) gives exactly the single-process result.  The forward is a stand-in function of the
batch (this test covers the host side only; GPU parity of the forward lives in tests/test_gpu_*.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qa_tiger_b200 import evalloop as EL


class TinySet:
    """Same interface as SyntheticAVQA, tiny tensors (CPU friendly)."""

    def __init__(self, n):
        self.n = n

    def batch(self, s, e, device):
        idx = torch.arange(s, e)
        g = torch.Generator().manual_seed(5)
        table = torch.randn(self.n, 42, generator=g)
        return dict(feat=table[idx], label=(idx * 7) % 42, qtype_label=(idx * 5) % 9)


def fake_forward(b):
    # logits depend on the LOCAL batch (like the real model): add the batch mean
    return b["feat"] + b["feat"].mean(dim=0, keepdim=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, bs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = EL.evaluate(fake_forward, TinySet(n), torch.device("cpu"), batch_size=bs, rank=rank, world=world)
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def _expected(n, bs, world):
    """Single-process evaluation of exactly the local batches the ranks use."""
    ds = TinySet(n)
    preds, counters = [], torch.zeros(20, dtype=torch.int64)
    for r in range(world):
        for (s, e) in EL.local_batches(n, r, world, bs):
            b = ds.batch(s, e, "cpu")
            p = fake_forward(b).argmax(1)
            preds.append(p)
            EL.pack_counters(p, b["label"], b["qtype_label"], counters)
    return torch.cat(preds), counters


@pytest.mark.parametrize("n,bs", [(101, 16), (64, 16), (7, 16), (33, 5)])
def test_two_ranks_match_single_process(tmp_path, n, bs):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, bs, out), nprocs=2, join=True)
    res = torch.load(out, weights_only=False)
    preds, counters = _expected(n, bs, 2)
    assert torch.equal(res["predictions"], preds)
    assert res["total"] == n == int(counters[1]) and res["correct"] == int(counters[0])
    assert res["tot"] == counters[2:11].tolist() and res["correct_by_type"] == counters[11:].tolist()
    assert sum(res["tot"]) == n


def test_sharding_is_contiguous_and_covers_everything():
    for n, w, bs in [(9129, 8, 256), (9129, 4, 256), (10, 8, 256), (0, 2, 4), (512, 2, 256)]:
        seen = []
        for r in range(w):
            lo, hi = EL.shard_range(n, r, w)
            bl = EL.local_batches(n, r, w, bs)
            assert all(e - s <= bs for s, e in bl)
            if bl:
                assert bl[0][0] == lo and bl[-1][1] == hi
            seen += [i for s, e in bl for i in range(s, e)]
        assert seen == list(range(n))


def test_report_format_matches_reference_lines():
    res = dict(acc=50.0, correct=9, total=18, tot=[2] * 9, correct_by_type=[1] * 9)
    lines = EL.format_report(res)
    assert lines[0] == "Test           Audio/Counting accuracy: 50.00(1/2)"
    assert lines[-1] == "Test                Total avg accuracy: 50.00(9/18)"
    assert len(lines) == 9 + 3 + 1


def test_synthetic_set_is_sharding_invariant():
    ds = EL.SyntheticAVQA(n=20, seed=3, T=4, P=2)
    a = ds.batch(3, 9, torch.device("cpu"))
    b = ds.batch(5, 7, torch.device("cpu"))
    for k in a:
        assert torch.equal(a[k][2:4], b[k]), k
    pool = EL.SyntheticAVQA(n=200, seed=3, T=4, P=2, n_unique_questions=33)
    q = torch.stack([pool._tokens(i) for i in range(200)])
    assert len(torch.unique(q, dim=0)) <= 33
