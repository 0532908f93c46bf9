"""CPU: host-side logic of the input pipeline (qa-tiger_b200/datapipe.py): the MUSIC-AVQA-shaped question
list, the per-video row layout, the device cache's slot / LRU / lock bookkeeping.  No compute calls (the gather kernel and
the streaming loader are covered by tests/test_gpu_datapipe.py)."""
import collections

import pytest
import torch

from qa_tiger_b200 import datapipe as dp


def test_question_set_has_the_test_sets_shape():
    q = dp.QuestionSet.synthetic_music_avqa()
    assert q.n == 9129 and q.n_videos == 6399
    mult = collections.Counter(collections.Counter(q.video.tolist()).values())
    assert dict(mult) == dp.MUSIC_AVQA_TEST_MULTIPLICITY
    assert len({bytes(r.numpy().tobytes()) for r in q.tokens}) == 33
    # ids are dense and numbered in first-use order: a store built as range(n_videos) has its rows in copy order
    seen = -1
    for v in q.video.tolist():
        assert v <= seen + 1
        seen = max(seen, v)
    # the same list on every call (sharding must see identical samples)
    q2 = dp.QuestionSet.synthetic_music_avqa()
    assert torch.equal(q.video, q2.video) and torch.equal(q.tokens, q2.tokens) and torch.equal(q.label, q2.label)
    small = dp.QuestionSet.synthetic_music_avqa(n=300, n_unique_questions=0)
    assert small.n == 300 and small.n_videos <= 300 and len({r.numpy().tobytes() for r in small.tokens}) == 300


def test_row_layout_views_are_views():
    L = dp.RowLayout(T=4, P=2, audio_dim=8, video_dim=8, patch_dim=16, patch_dtype=torch.float16)
    assert L.nbytes == (128, 128, 256) and L.offset == (0, 128, 256) and L.row_bytes == 512
    rows = torch.zeros(3, L.row_bytes, dtype=torch.uint8)
    a, v, p = L.views(rows)
    assert a.shape == (3, 4, 8) and v.shape == (3, 4, 8) and p.shape == (3, 4, 2, 16) and p.dtype == torch.float16
    a[1, 0, 0], v[2, 3, 7], p[0, 3, 1, 15] = 1.0, 2.0, 3.0
    assert rows[1, :4].tolist() == list(torch.tensor([1.0]).view(torch.uint8))
    a2, v2, p2 = L.views(rows[1:])               # a slice with a storage offset still maps the right bytes
    assert float(a2[0, 0, 0]) == 1.0 and float(v2[1, 3, 7]) == 2.0
    with pytest.raises(ValueError):
        dp.RowLayout(T=1, P=1, audio_dim=3, video_dim=8, patch_dim=16)
    L32 = dp.RowLayout(patch_dtype=torch.float32)
    assert L32.row_bytes == 60 * 128 * 4 + 60 * 768 * 4 + 60 * 14 * 1024 * 4


def test_synthetic_store_is_a_function_of_the_video_id():
    L = dp.RowLayout(T=4, P=2, audio_dim=8, video_dim=8, patch_dim=16)
    a = dp.FeatureStore.synthetic([0, 1, 2, 3], L)
    b = dp.FeatureStore.synthetic([3, 1], L)
    assert torch.equal(a.rows[3], b.rows[0]) and torch.equal(a.rows[1], b.rows[1])
    assert not torch.equal(a.rows[0], a.rows[1])
    with pytest.raises(ValueError):
        dp.FeatureStore(L, [1, 1])


def test_cache_slots_lru_and_locks():
    L = dp.RowLayout(T=4, P=2, audio_dim=8, video_dim=8, patch_dim=16)
    c = dp.DeviceFeatureCache(L, 3, torch.device("cpu"), allocate=False)
    slots, copy_in, ev = c.lookup([10, 11, 10])
    assert slots == [0, 1, 0] and copy_in == [(10, 0), (11, 1)] and not ev
    # batch still in flight: its slots cannot be evicted
    slots2, copy2, ev2 = c.lookup([12])
    assert slots2 == [2] and copy2 == [(12, 2)] and not ev2
    with pytest.raises(RuntimeError, match="capacity"):
        c.lookup([13])
    assert c.locked == {0: 2, 1: 1, 2: 1}                     # the failed lookup rolled its locks back
    c.release(slots)
    slots3, copy3, ev3 = c.lookup([13])                       # evicts the least recently used unlocked video: 11 (10 was
    assert ev3 and copy3 == [(13, 1)] and 11 not in c.slot_of  # touched again after 11)
    c.release(slots2); c.release(slots3)
    s4, cp4, _ = c.lookup([10])                               # a hit refreshes recency and copies nothing
    assert s4 == [0] and cp4 == []
    c.release(s4)
    s5, cp5, ev5 = c.lookup([14])                             # ... so 12 is the LRU victim now, not 10
    assert ev5 and cp5 == [(14, 2)] and 10 in c.slot_of
    st = c.stats()
    assert st["hits"] == 2 and st["misses"] == 5 and st["evictions"] == 2 and st["h2d_bytes"] == 5 * L.row_bytes
    c.clear()
    assert not c.slot_of and len(c.free) == 3


def test_loader_refuses_a_cpu_device():
    L = dp.RowLayout(T=4, P=2, audio_dim=8, video_dim=8, patch_dim=16)
    q = dp.QuestionSet.synthetic_music_avqa(n=20)
    st = dp.FeatureStore.synthetic(range(q.n_videos), L)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dp.StreamingLoader(st, q, torch.device("cpu"))


def test_failed_lookup_leaves_no_phantom_residents():
    L = dp.RowLayout(T=4, P=2, audio_dim=8, video_dim=8, patch_dim=16)
    c = dp.DeviceFeatureCache(L, 2, torch.device("cpu"), allocate=False)
    with pytest.raises(RuntimeError, match="capacity"):
        c.lookup([1, 2, 3])                                   # 1 and 2 got slots before 3 found none: both must be undone
    assert not c.slot_of and not c.locked and sorted(c.free) == [0, 1]
    slots, copy_in, _ = c.lookup([1, 2])
    assert sorted(slots) == [0, 1] and len(copy_in) == 2
