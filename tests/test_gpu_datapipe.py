"""GPU: the input pipeline and the eval loop (8f-N2) on the real model.
  * qat_gather_rows against torch indexing, bit for bit (strided cache rows, fp16 and fp32 patch stores, ragged n);
  * StreamingLoader (per-video device cache, prefetch, text tower on its own stream, memo) yields exactly the batches the
    reference-style DirectDataset does, cold and warm, with and without evictions;
  * evalloop.evaluate through the loader == the oracle evaluated on `local_batches` (identical predictions and counters)."""
import numpy as np
import pytest
import torch

from oracle import qat_oracle as O
from oracle import synth

from gpu_util import build_model, dev, relerr

pytestmark = pytest.mark.gpu


def _small(n=70, patch_dtype=torch.float16, n_unique=5, seed=1234):
    from qa_tiger_b200 import datapipe as dp
    q = dp.QuestionSet.synthetic_music_avqa(n=n, n_unique_questions=n_unique, seed=seed)
    L = dp.RowLayout(patch_dtype=patch_dtype)
    st = dp.FeatureStore.synthetic(range(q.n_videos), L, seed=seed, device=dev())
    return dp, q, L, st


@pytest.mark.parametrize("patch_dtype", [torch.float16, torch.float32])
def test_gather_rows_matches_indexing(patch_dtype):
    from qa_tiger_b200 import datapipe as dp
    L = dp.RowLayout(T=6, P=3, audio_dim=12, video_dim=20, patch_dim=24, patch_dtype=patch_dtype)
    g = torch.Generator().manual_seed(5)
    cache = torch.randint(0, 255, (37, L.row_bytes), dtype=torch.uint8, generator=g).to(dev())
    for n in (1, 5, 64):
        slots = torch.randint(0, 37, (n,), generator=g).to(torch.int32).to(dev())
        out = tuple(torch.zeros((n,) + shp, dtype=dt, device=dev()) for shp, dt in zip(L.shapes, L.dtypes))
        dp.gather_rows(cache, L, slots, out)
        ref = L.views(cache[slots.long()])
        for o, r in zip(out, ref):
            assert torch.equal(o.view(torch.uint8), r.contiguous().view(torch.uint8))
    # the real row: 1.9 MB, many CTAs per row, a vector count that is not a multiple of the CTA pass
    L = dp.RowLayout()
    cache = torch.randint(0, 255, (9, L.row_bytes), dtype=torch.uint8, generator=g).to(dev())
    slots = torch.tensor([8, 0, 3, 3, 7], dtype=torch.int32, device=dev())
    out = tuple(torch.zeros((5,) + shp, dtype=dt, device=dev()) for shp, dt in zip(L.shapes, L.dtypes))
    dp.gather_rows(cache, L, slots, out)
    for o, r in zip(out, L.views(cache[slots.long()])):
        assert torch.equal(o.view(torch.uint8), r.contiguous().view(torch.uint8))
    # a slot outside the cache poisons the row (NaN bit patterns) instead of reading out of bounds
    bad = torch.tensor([1, 9], dtype=torch.int32, device=dev())
    out = tuple(torch.zeros((2,) + shp, dtype=dt, device=dev()) for shp, dt in zip(L.shapes, L.dtypes))
    dp.gather_rows(cache, L, bad, out)
    assert torch.isnan(out[0][1]).all() and torch.isnan(out[2][1].float()).all()
    assert torch.equal(out[0][0].view(torch.uint8), L.views(cache[1:2])[0][0].contiguous().view(torch.uint8))


@pytest.mark.parametrize("capacity", [None, 40])
def test_streaming_loader_equals_direct_batches(capacity):
    """capacity None = every video of the shard stays resident; 40 = below the working set, LRU evictions on every batch"""
    dp, q, L, st = _small(n=70)
    direct = dp.DirectDataset(st, q)
    loader = dp.StreamingLoader(st, q, dev(), batch_size=16, capacity=capacity)
    for rep in range(2):                                   # cold, then warm (second pass: everything resident)
        seen = 0
        for s, e, b in loader.iter_batches():
            ref = direct.batch(s, e, dev())
            for k in ("quest", "audio", "video", "patch", "label", "qtype_label"):
                assert torch.equal(b[k], ref[k]), (rep, s, k)
            seen += e - s
        assert seen == q.n
    stats = loader.cache.stats()
    if capacity is None:
        assert stats["misses"] == q.n_videos and stats["evictions"] == 0       # every video crossed the link exactly once
        assert stats["hits"] == 2 * q.n - q.n_videos
    else:
        assert stats["evictions"] > 0 and stats["resident"] <= 40


def test_streaming_eval_equals_oracle_on_local_batches():
    """SURVEY 8f-N2 + N3 on hardware: evaluate() through the streaming loader (text tower on the loader's stream, memo on,
    on-device counters) against the oracle run batch by batch on `local_batches`: identical predictions and counters."""
    from qa_tiger_b200 import evalloop as EL
    dp, q, L, st = _small(n=41, patch_dtype=torch.float32, n_unique=4)
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "fp32")
    model.quest_encoder.memo = True
    loader = dp.StreamingLoader(st, q, dev(), batch_size=16, encoder=model.quest_encoder)
    res = EL.evaluate(lambda b: model(b)["out"], loader, dev(), 16, 0, 1)
    a, v, p = st.typed()
    preds, logits_ref = [], []
    for s, e in EL.local_batches(q.n, 0, 1, 16):
        rows = torch.tensor([st.index[x] for x in q.video[s:e].tolist()])
        with torch.no_grad():
            out = O.qa_tiger_forward(sd, dict(quest=q.tokens[s:e].clone(), audio=a[rows].clone(), video=v[rows].clone(), patch=p[rows].clone()))
        preds.append(out.argmax(1))
        logits_ref.append(out)
    preds = torch.cat(preds)
    assert res["predictions"].tolist() == preds.tolist()
    hit = preds == q.label
    assert res["correct"] == int(hit.sum()) and res["total"] == q.n
    assert res["tot"] == torch.bincount(q.qtype, minlength=9).tolist()
    assert res["correct_by_type"] == torch.bincount(q.qtype[hit], minlength=9).tolist()
    # second pass: every video and question resident, same answers; the tower ran on 4 questions in total
    res2 = EL.evaluate(lambda b: model(b)["out"], loader, dev(), 16, 0, 1)
    assert res2["predictions"].tolist() == preds.tolist()
    assert len(model.quest_encoder._memo_state["slots"]) == 4
    assert loader.cache.stats()["misses"] == q.n_videos


def test_streaming_eval_bf16_fp16_store_within_gate():
    """the shipping configuration: fp16 patch store, bf16 mode, tower through the loader -- logits within the 1e-3 gate of
    the oracle on the same (fp16-representable) features, identical answers"""
    from qa_tiger_b200 import evalloop as EL
    dp, q, L, st = _small(n=24, patch_dtype=torch.float16, n_unique=3)
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "bf16")
    model.quest_encoder.memo = True
    loader = dp.StreamingLoader(st, q, dev(), batch_size=24, encoder=model.quest_encoder)
    outs = []
    res = EL.evaluate(lambda b: outs.append(model(b)["out"]) or outs[-1], loader, dev(), 24, 0, 1)
    a, v, p = st.typed()
    with torch.no_grad():
        ref = O.qa_tiger_forward(sd, dict(quest=q.tokens.clone(), audio=a.clone()[q.video], video=v.clone()[q.video], patch=p.float()[q.video]))
    assert relerr(outs[0], ref) < 1e-3
    assert res["predictions"].tolist() == ref.argmax(1).tolist()


def test_memo_is_exact_and_capped():
    """ADVICE: the memo keys on the exact token row; beyond `memo_max_entries` questions are encoded but not remembered"""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "bf16")
    enc = model.quest_encoder
    toks = synth.make_batch(3, 6, 60, with_patch=False)["quest"]
    with torch.no_grad():
        q_ref, w_ref = enc(toks.to(dev()))
        enc.memo, enc.memo_max_entries = True, 4
        q1, w1 = enc(toks.to(dev()), text_host=toks)                       # 6 distinct questions, room for 4
        assert len(enc._memo_state["slots"]) == 4
        assert torch.equal(q1, q_ref) and torch.equal(w1, w_ref)
        q2, w2 = enc(toks[[5, 0, 4, 0]].to(dev()))                         # hits, an uncached question twice, no text_host
        assert torch.equal(q2, q_ref[[5, 0, 4, 0]]) and torch.equal(w2, w_ref[[5, 0, 4, 0]])
        assert len(enc._memo_state["slots"]) == 4
        bad = toks.clone()
        bad[2, 3] = 49408
        with pytest.raises(RuntimeError, match="token id outside"):
            enc(bad.to(dev()), text_host=bad)
    enc.memo = False


def test_bad_token_id_poisons_instead_of_clamping():
    """nn.Embedding raises on an out-of-range id; the native tower (which cannot raise from the device) returns NaN for that
    question, never the encoding of a clamped id"""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "bf16")
    toks = synth.make_batch(3, 3, 60, with_patch=False)["quest"].clone()
    toks[1, 2] = -7
    with torch.no_grad():
        q, w = model.quest_encoder(toks.to(dev()))
    assert torch.isnan(q[1]).all() and not torch.isnan(q[0]).any() and not torch.isnan(q[2]).any()
