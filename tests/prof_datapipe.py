"""Profiling aid (not a test): where the cold streaming pass spends its time.
  python tests/prof_datapipe.py [n_questions]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from qa_tiger_b200 import datapipe as dp, evalloop as EL  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import build_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9129
dev = torch.device("cuda:0")
qs = dp.QuestionSet.synthetic_music_avqa(n=n)
L = dp.RowLayout()
t0 = time.perf_counter()
store = dp.FeatureStore.synthetic(range(qs.n_videos), L, device=dev)
torch.cuda.synchronize()
print(f"store {store.nbytes / 1e9:.2f} GB pinned={store.rows.is_pinned()} built in {time.perf_counter() - t0:.1f}s")


def timed(fn, name, nbytes=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {dt * 1e3:.1f} ms (host issue {t_issue * 1e3:.1f} ms)" + (f"  {nbytes / dt / 1e9:.1f} GB/s" if nbytes else ""))
    return dt


# (1) raw H2D of the whole store in 200-row slices
dst = torch.empty_like(store.rows, device=dev)
cs = torch.cuda.Stream()


def raw(step):
    with torch.cuda.stream(cs):
        for r in range(0, store.rows.shape[0], step):
            dst[r:r + step].copy_(store.rows[r:r + step], non_blocking=True)
    torch.cuda.current_stream().wait_stream(cs)


for step in (200, 200, 1, 4096):
    timed(lambda: raw(step), f"raw H2D, {step} rows per copy", store.nbytes)
del dst

sd = synth.make_state_dict(713, num_experts=7)
model = build_model(sd, 7, 7, "bf16")
model.quest_encoder.memo = True
loader = dp.StreamingLoader(store, qs, dev, batch_size=256, encoder=model.quest_encoder)
fwd = lambda b: model(b)["out"]
with torch.no_grad():
    EL.evaluate(fwd, loader, dev, 256)

    def cold():
        loader.cache.clear()
        model.quest_encoder.__dict__.pop("_memo_state", None)

    def loader_only():
        for s, e, b in loader.iter_batches():
            pass
    cold()
    timed(loader_only, "cold loader only (copies + tower + gather, no forward)", store.nbytes)
    timed(loader_only, "warm loader only")
    cold()
    timed(lambda: EL.evaluate(fwd, loader, dev, 256), "cold evaluate", store.nbytes)
    timed(lambda: EL.evaluate(fwd, loader, dev, 256), "warm evaluate")
    # per-batch host time of the prefetch
    cold()
    ts = []
    plan = EL.local_batches(qs.n, 0, 1, 256)
    for i, (s, e) in enumerate(plan[:6]):
        t0 = time.perf_counter()
        rec = loader._prefetch(i % 2, s, e)
        ts.append((time.perf_counter() - t0) * 1e3)
        loader.cache.release(rec["slots"])
    torch.cuda.synchronize()
    print("prefetch host ms per batch:", [f"{t:.2f}" for t in ts])
