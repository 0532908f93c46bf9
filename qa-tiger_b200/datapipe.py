"""Input pipeline of the eval loop: the B200 counterpart of /root/reference/src/dataset.py:132-159
(`AVQA_dataset.load_samples`: three `np.load`s per QUESTION) and src/trainutils.py:223-236 (`get_items`: every batch
`.to(device).float()`), feeding the test loop of trainutils.py:411-423.

What changes against the reference pipeline:
  * features are kept per VIDEO, not per question (MUSIC-AVQA test: 9 129 questions over 6 399 videos): `FeatureStore`
    holds one pinned host row per video = audio | video | patch back to back, so a video is ONE host->device copy;
  * `DeviceFeatureCache` keeps every video that has crossed the host link in HBM (the whole fp16 test set is 12.4 GB of
    a B200's 180 GB; LRU eviction only when the capacity is set below the working set), so a video crosses PCIe once
    per eval -- and never again in the following evals (the reference re-evaluates the same set every epoch, train.py);
  * the patch tokens may be stored as fp16 (`patch_dtype=torch.float16`): half the bytes on disk, in host memory, over the
    link and in HBM; the reference's own `get_items` (`.to(device).float()`) accepts such a store unchanged, and this
    path's patch_proj consumes it directly as one kind::f16 tensor-core pass (QatPatchArgs.patch_f16);
  * `StreamingLoader` overlaps, for batch i+1 while batch i computes: the host->device copies of its unseen videos (copy
    stream, straight from the pinned store), its token / label / slot vectors, and the CLIP text tower on its unseen
    questions (tower stream; keyed on the exact token row, models/encoders.py).  The batch itself is assembled in HBM by
    ONE launch of `qat_gather_rows` (csrc/datapipe.cuh) -- the only kernel of this module.
No CPU path: the loader needs CUDA (host-side bookkeeping -- slots, LRU, the batch plan -- is plain Python and is what the
`-m "not gpu"` tests cover).
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi as cabi
from .evalloop import N_QTYPES, local_batches, shard_range

# multiplicity histogram of music_avqa_test.json (questions per video): 4083 videos with one question, 1928 with two, ...
# -> 6 399 videos, 9 129 questions; repeats of a video are spread over the file (3 adjacent pairs in 9 129), which is why
# a per-batch dedup would not help and a cache does
MUSIC_AVQA_TEST_MULTIPLICITY = {1: 4083, 2: 1928, 3: 364, 4: 22, 5: 2}


class RowLayout:
    """Byte layout of one video's row: audio [T,Da] f32 | video [T,Dv] f32 | patch [T,P,Dp] f16|f32, 16-byte aligned parts."""

    def __init__(self, T: int = 60, P: int = 14, audio_dim: int = 128, video_dim: int = 768, patch_dim: int = 1024,
                 patch_dtype: torch.dtype = torch.float16):
        if patch_dtype not in (torch.float16, torch.float32):
            raise ValueError("patch_dtype must be torch.float16 or torch.float32")
        self.T, self.P, self.audio_dim, self.video_dim, self.patch_dim, self.patch_dtype = T, P, audio_dim, video_dim, patch_dim, patch_dtype
        pe = 2 if patch_dtype == torch.float16 else 4
        self.nbytes = (T * audio_dim * 4, T * video_dim * 4, T * P * patch_dim * pe)
        if any(b % 16 for b in self.nbytes):
            raise ValueError(f"feature rows must be multiples of 16 bytes, got {self.nbytes}")
        self.offset = (0, self.nbytes[0], self.nbytes[0] + self.nbytes[1])
        self.row_bytes = sum(self.nbytes)
        self.shapes = ((T, audio_dim), (T, video_dim), (T, P, patch_dim))
        self.dtypes = (torch.float32, torch.float32, patch_dtype)

    def views(self, rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """typed (audio, video, patch) views of a [n, row_bytes] uint8 tensor (no copy)"""
        n = rows.shape[0]
        out = []
        for off, nb, shp, dt in zip(self.offset, self.nbytes, self.shapes, self.dtypes):
            es = dt.itemsize
            out.append(rows.view(dt)[:, off // es:(off + nb) // es].view((n,) + shp))     # column slice, then split: still a view
        return tuple(out)


class FeatureStore:
    """Host-side feature store: one pinned row per video (`RowLayout`).  `index[vid]` -> row."""

    def __init__(self, layout: RowLayout, video_ids: Sequence[int], pin: bool = True):
        self.layout = layout
        self.video_ids = [int(v) for v in video_ids]
        self.index = {v: i for i, v in enumerate(self.video_ids)}
        if len(self.index) != len(self.video_ids):
            raise ValueError("FeatureStore: duplicate video ids")
        pin = pin and torch.cuda.is_available()
        self.rows = torch.empty(len(self.video_ids), layout.row_bytes, dtype=torch.uint8, pin_memory=pin)
        self.pinned = pin

    @property
    def nbytes(self) -> int:
        return self.rows.numel()

    def typed(self):
        return self.layout.views(self.rows)

    @classmethod
    def synthetic(cls, video_ids: Sequence[int], layout: Optional[RowLayout] = None, seed: int = 1234,
                  device: Optional[torch.device] = None, chunk: int = 128) -> "FeatureStore":
        """N(0,1) features, a pure function of (seed, video id): every sharding sees identical videos.  Generated on
        `device` (a GPU when given: 12 GB of host-side randn would take minutes) and copied into the pinned rows."""
        layout = layout or RowLayout()
        st = cls(layout, video_ids)
        gen_dev = device if device is not None else torch.device("cpu")
        g = torch.Generator(device=gen_dev)
        for c0 in range(0, len(st.video_ids), chunk):
            ids = st.video_ids[c0:c0 + chunk]
            buf = torch.empty(len(ids), layout.row_bytes, dtype=torch.uint8, device=gen_dev)
            a, v, p = layout.views(buf)
            for j, vid in enumerate(ids):
                g.manual_seed(seed * 7919 + vid)
                a[j].normal_(generator=g)
                v[j].normal_(generator=g)
                if layout.patch_dtype == torch.float16:
                    p[j].copy_(torch.empty(layout.shapes[2], device=gen_dev).normal_(generator=g))
                else:
                    p[j].normal_(generator=g)
            st.rows[c0:c0 + len(ids)].copy_(buf)
        return st

    @classmethod
    def from_npy(cls, names: Sequence[str], audio_dir: str, video_dir: str, patch_dir: str, layout: Optional[RowLayout] = None,
                 sample_rate: int = 1) -> "FeatureStore":
        """The reference's on-disk format: `<dir>/<video name>.npy` per modality (dataset.py:136-157); video ids are the
        positions in `names`.  Loaded ONCE per video."""
        layout = layout or RowLayout()
        st = cls(layout, range(len(names)))
        a, v, p = st.typed()
        for i, name in enumerate(names):
            a[i].copy_(torch.from_numpy(np.load(os.path.join(audio_dir, f"{name}.npy"))).float())
            v[i].copy_(torch.from_numpy(np.load(os.path.join(video_dir, f"{name}.npy")))[::sample_rate].float())
            p[i].copy_(torch.from_numpy(np.load(os.path.join(patch_dir, f"{name}.npy")))[::sample_rate])
        return st


class QuestionSet:
    """The annotation side of AVQA_dataset (dataset.py:107-131): per question its video id, CLIP token ids, answer label
    and question-type id; host tensors (pinned when CUDA is there: batches are contiguous slices, copied as they are)."""

    def __init__(self, video: torch.Tensor, tokens: torch.Tensor, label: torch.Tensor, qtype: torch.Tensor):
        n = video.numel()
        if tokens.shape[0] != n or label.numel() != n or qtype.numel() != n:
            raise ValueError("QuestionSet: per-question arrays disagree in length")
        pin = torch.cuda.is_available()
        mk = (lambda t: t.contiguous().pin_memory()) if pin else (lambda t: t.contiguous())
        self.video = video.to(torch.int64).contiguous()
        self.tokens = mk(tokens.to(torch.int32))
        self.label = mk(label.to(torch.int64).reshape(-1))
        self.qtype = mk(qtype.to(torch.int64).reshape(-1))
        self.n = n

    def videos_of(self, lo: int, hi: int) -> List[int]:
        """distinct video ids of questions [lo, hi), in first-use order"""
        return list(OrderedDict.fromkeys(self.video[lo:hi].tolist()))

    @classmethod
    def synthetic_music_avqa(cls, n: int = 9129, n_unique_questions: int = 33, seed: int = 1234,
                             multiplicity: Optional[Dict[int, int]] = None) -> "QuestionSet":
        """A MUSIC-AVQA-test-shaped question list: the questions-per-video histogram of the real annotation file
        (`MUSIC_AVQA_TEST_MULTIPLICITY`, scaled to `n`), repeats scattered over the list by a seeded permutation, token rows
        drawn from a pool of `n_unique_questions` strings (33 in the real set; 0 = all distinct)."""
        mult = multiplicity or MUSIC_AVQA_TEST_MULTIPLICITY
        total = sum(k * c for k, c in mult.items())
        vids: List[int] = []
        nxt = 0
        for k in sorted(mult):
            cnt = mult[k] if n == total else max(1, round(mult[k] * n / total))
            for _ in range(cnt):
                vids += [nxt] * k
                nxt += 1
        while len(vids) < n:                     # rounding remainder: single-question videos
            vids.append(nxt)
            nxt += 1
        g = torch.Generator().manual_seed(seed)
        vids_t = torch.tensor(vids, dtype=torch.int64)[torch.randperm(len(vids), generator=g)][:n]
        # renumber in first-use order so ids are dense even after the truncation
        remap: Dict[int, int] = {}
        dense = [remap.setdefault(v, len(remap)) for v in vids_t.tolist()]
        tokens = torch.zeros(n, 77, dtype=torch.int32)
        pool: Dict[int, torch.Tensor] = {}
        for i in range(n):
            key = i if n_unique_questions <= 0 else (i * 2654435761) % n_unique_questions
            row = pool.get(key)
            if row is None:
                gq = torch.Generator().manual_seed(seed * 1000003 + key)
                row = torch.zeros(77, dtype=torch.int32)
                L = int(torch.randint(5, 20, (1,), generator=gq))
                row[0] = 49406
                row[1:1 + L] = torch.randint(1, 49406, (L,), generator=gq, dtype=torch.int32)
                row[1 + L] = 49407
                pool[key] = row
            tokens[i] = row
        idx = torch.arange(n, dtype=torch.int64)
        h = (idx * 1103515245 + seed) & 0x7FFFFFFF
        return cls(torch.tensor(dense), tokens, h % 42, (h // 42) % N_QTYPES)

    @property
    def n_videos(self) -> int:
        return int(self.video.max()) + 1 if self.n else 0


class DeviceFeatureCache:
    """HBM-resident per-video rows, keyed by video id.  Host-side bookkeeping only decides WHICH slot a video gets; the
    bytes move in `StreamingLoader`.  LRU eviction; slots of batches that are still in flight are locked."""

    def __init__(self, layout: RowLayout, capacity: int, device: torch.device, allocate: bool = True):
        if capacity < 1:
            raise ValueError("DeviceFeatureCache: capacity must be >= 1")
        self.layout, self.capacity, self.device = layout, int(capacity), device
        self.rows = torch.empty(self.capacity, layout.row_bytes, dtype=torch.uint8, device=device) if allocate else None
        self.slot_of: "OrderedDict[int, int]" = OrderedDict()        # vid -> slot, least recently used first
        self.free: List[int] = list(range(self.capacity - 1, -1, -1))
        self.locked: Dict[int, int] = {}                             # slot -> lock count
        self.hits = self.misses = self.evictions = 0

    def lookup(self, video_ids: Iterable[int]) -> Tuple[List[int], List[Tuple[int, int]], bool]:
        """slots of `video_ids` (one per id, duplicates allowed) + the (vid, slot) pairs that must be copied in + whether a
        resident video was evicted to make room.  Every returned slot is locked until `release()`."""
        slots: List[int] = []
        copy_in: List[Tuple[int, int]] = []
        evicted = False
        taken: List[int] = []
        hits = misses = 0
        for vid in video_ids:
            s = self.slot_of.get(vid)
            if s is not None:
                self.slot_of.move_to_end(vid)
                hits += 1
            else:
                misses += 1
                if self.free:
                    s = self.free.pop()
                else:
                    victim = next((v for v, sl in self.slot_of.items() if not self.locked.get(sl)), None)
                    if victim is None:
                        for t in taken:                          # roll this call back before raising: its locks, and the
                            self._unlock(t)                      # videos it had given slots to (they were never copied in)
                        for v_, s_ in copy_in:
                            del self.slot_of[v_]
                            self.free.append(s_)
                        raise RuntimeError(f"DeviceFeatureCache: capacity {self.capacity} is below the videos of the batches in flight")
                    s = self.slot_of.pop(victim)
                    self.evictions += 1
                    evicted = True
                self.slot_of[vid] = s
                copy_in.append((vid, s))
            self.locked[s] = self.locked.get(s, 0) + 1
            taken.append(s)
            slots.append(s)
        self.hits += hits
        self.misses += misses
        return slots, copy_in, evicted

    def _unlock(self, s: int) -> None:
        c = self.locked.get(s, 0) - 1
        if c <= 0:
            self.locked.pop(s, None)
        else:
            self.locked[s] = c

    def release(self, slots: Iterable[int]) -> None:
        for s in slots:
            self._unlock(s)

    def clear(self) -> None:
        self.slot_of.clear()
        self.free = list(range(self.capacity - 1, -1, -1))
        self.locked.clear()
        self.hits = self.misses = self.evictions = 0

    def stats(self) -> Dict[str, int]:
        return dict(hits=self.hits, misses=self.misses, evictions=self.evictions, resident=len(self.slot_of), capacity=self.capacity,
                    h2d_bytes=self.misses * self.layout.row_bytes)


def gather_rows(cache_rows: torch.Tensor, layout: RowLayout, slots_dev: torch.Tensor, out: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]):
    """out[k][i] = part k of cache row slots_dev[i]  (one launch of qat_gather_rows)"""
    n = slots_dev.numel()
    a = cabi.GatherArgs()
    a.n, a.n_tensors, a.n_slots, a.stream = n, 3, cache_rows.shape[0], cabi.stream_ptr(cache_rows.device)
    a.slots = slots_dev.data_ptr()
    base = cache_rows.data_ptr()
    for k in range(3):
        a.src[k] = base + layout.offset[k]
        a.dst[k] = out[k].data_ptr()
        a.row_bytes[k] = layout.nbytes[k]
        a.src_stride[k] = cache_rows.stride(0)
    with torch.cuda.device(cache_rows.device):
        cabi.check(cabi.lib().qat_gather_rows(C.byref(a)), "qat_gather_rows")


class StreamingLoader:
    """Iterates this rank's local batches (contiguous shard, in order -- the partition is part of the parity contract,
    SURVEY 8e) and yields `reshaped_data` dicts on `device`, one batch ahead of the consumer.

    `encoder`: the model's CLIP_TEncoder; when given, the text tower runs here (tower stream, memo on) and the batch carries
    `quest_feat = (quest768, words768)`, which QA_TIGER.forward consumes instead of encoding `quest` itself.
    Use as `dataset` of `evalloop.evaluate` (it provides `iter_batches`)."""

    def __init__(self, store: FeatureStore, questions: QuestionSet, device: torch.device, batch_size: int = 256, rank: int = 0,
                 world: int = 1, cache: Optional[DeviceFeatureCache] = None, capacity: Optional[int] = None, encoder=None):
        if device.type != "cuda":
            raise RuntimeError("StreamingLoader needs a CUDA device (qa-tiger_b200 has no CPU path)")
        cabi.lib()
        self.store, self.q, self.device, self.batch_size, self.rank, self.world = store, questions, device, batch_size, rank, world
        self.n = questions.n
        self.layout = store.layout
        lo, hi = shard_range(self.n, rank, world)
        need = len(questions.videos_of(lo, hi))
        if cache is None:
            cache = DeviceFeatureCache(self.layout, capacity if capacity is not None else max(1, need), device)
        self.cache = cache
        self.encoder = encoder
        self.copy_stream = torch.cuda.Stream(device)
        self.tower_stream = torch.cuda.Stream(device)
        B = batch_size
        # two sets of per-batch vectors (the batch being consumed and the one being prefetched)
        self._slots_host = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._slots_dev = [torch.empty(B, dtype=torch.int32, device=device) for _ in range(2)]
        self._tok_dev = [torch.empty(B, questions.tokens.shape[1], dtype=torch.int32, device=device) for _ in range(2)]
        self._lab_dev = [torch.empty(B, dtype=torch.int64, device=device) for _ in range(2)]
        self._qt_dev = [torch.empty(B, dtype=torch.int64, device=device) for _ in range(2)]
        # ONE set of batch tensors: gather(i+1) follows forward(i) on the consumer's stream
        L = self.layout
        self._batch = tuple(torch.empty((B,) + shp, dtype=dt, device=device) for shp, dt in zip(L.shapes, L.dtypes))
        self._gather_done: Optional[torch.cuda.Event] = None
        self._copied: List[Optional[torch.cuda.Event]] = [None, None]
        self.h2d_bytes = 0

    def _prefetch(self, k: int, s: int, e: int):
        """issue everything batch [s, e) needs ahead of time; returns the in-flight record"""
        q, n = self.q, e - s
        vids = q.video[s:e].tolist()
        missing = [v for v in set(vids) if v not in self.store.index]
        if missing:
            raise RuntimeError(f"StreamingLoader: videos {missing[:4]}... are not in the feature store")
        slots, copy_in, evicted = self.cache.lookup(vids)
        if self._copied[k] is not None:
            self._copied[k].synchronize()                 # the pinned staging vector of set k is free again (and the host never
                                                          # runs more than two batches ahead of the copy engine)
        sh = self._slots_host[k]
        sh[:n] = torch.tensor(slots, dtype=torch.int32)
        # unseen videos straight from the pinned store, one copy per RUN of consecutive (store row, slot) pairs: a cold cache
        # hands out slots in order and a store built in first-use order (`QuestionSet.videos_of`) has its rows in that order
        # too, so a whole batch's misses are typically a handful of multi-megabyte copies
        pairs = sorted((slot, self.store.index[vid]) for vid, slot in copy_in)
        runs: List[List[int]] = []
        for slot, row in pairs:
            if runs and runs[-1][0] + runs[-1][2] == slot and runs[-1][1] + runs[-1][2] == row:
                runs[-1][2] += 1
            else:
                runs.append([slot, row, 1])
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            # the small per-question vectors first (their own event: the text tower only needs the tokens), then the videos
            self._slots_dev[k][:n].copy_(sh[:n], non_blocking=True)
            self._tok_dev[k][:n].copy_(q.tokens[s:e], non_blocking=True)
            self._lab_dev[k][:n].copy_(q.label[s:e], non_blocking=True)
            self._qt_dev[k][:n].copy_(q.qtype[s:e], non_blocking=True)
            vec_copied = torch.cuda.Event()
            vec_copied.record(cs)
            if evicted and self._gather_done is not None:
                cs.wait_event(self._gather_done)          # an overwritten slot may still be read by the last issued gather
            for slot, row, cnt in runs:
                self.cache.rows[slot:slot + cnt].copy_(self.store.rows[row:row + cnt], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(cs)
        self._copied[k] = copied
        nbytes = len(copy_in) * self.layout.row_bytes + n * (4 + q.tokens.shape[1] * 4 + 16)
        self.h2d_bytes += nbytes
        feat = towered = None
        if self.encoder is not None:
            ts = self.tower_stream
            with torch.cuda.stream(ts):
                ts.wait_event(vec_copied)
                feat = self.encoder(self._tok_dev[k][:n], text_host=q.tokens[s:e])
                towered = torch.cuda.Event()
                towered.record(ts)
        return dict(k=k, s=s, e=e, slots=slots, copied=copied, feat=feat, towered=towered, h2d_bytes=nbytes)

    def iter_batches(self, rank: Optional[int] = None, world: Optional[int] = None, batch_size: Optional[int] = None,
                     device: Optional[torch.device] = None) -> Iterator[Tuple[int, int, Dict[str, torch.Tensor]]]:
        if (rank, world, batch_size) != (None, None, None) and (rank, world, batch_size) != (self.rank, self.world, self.batch_size):
            raise RuntimeError("StreamingLoader was built for a different (rank, world, batch_size)")
        plan = local_batches(self.n, self.rank, self.world, self.batch_size)
        if not plan:
            return
        main = torch.cuda.current_stream(self.device)
        # the prefetch of batch 0 may reuse vectors the previous pass's last gather / forward still reads
        self.copy_stream.wait_stream(main)
        self.tower_stream.wait_stream(main)
        nxt = self._prefetch(0, *plan[0])
        for i, (s, e) in enumerate(plan):
            cur = nxt
            nxt = self._prefetch((i + 1) % 2, *plan[i + 1]) if i + 1 < len(plan) else None
            n, k = e - s, cur["k"]
            main.wait_event(cur["copied"])
            out = tuple(t[:n] for t in self._batch)
            gather_rows(self.cache.rows, self.layout, self._slots_dev[k][:n], out)
            self._gather_done = torch.cuda.Event()
            self._gather_done.record(main)
            batch = dict(quest=self._tok_dev[k][:n], audio=out[0], video=out[1], patch=out[2], label=self._lab_dev[k][:n],
                         qtype_label=self._qt_dev[k][:n], h2d_bytes=cur["h2d_bytes"])
            if cur["feat"] is not None:
                main.wait_event(cur["towered"])
                for t in cur["feat"]:
                    t.record_stream(main)                 # allocated on the tower stream, read on the consumer's
                batch["quest_feat"] = cur["feat"]
            self.cache.release(cur["slots"])              # the gather is issued: eviction waits on its event, not on the lock
            yield s, e, batch
            # the consumer's work on this batch is now in `main`; the vectors of set k may be overwritten once it is done
            done = torch.cuda.Event()
            done.record(main)
            self.copy_stream.wait_event(done)
            self.tower_stream.wait_event(done)


class DirectDataset:
    """The reference's behaviour on the same store, for A/B and parity: every batch copies every question's features from
    the host (no per-video cache, no prefetch), `dataset.batch(s, e, device)` protocol of evalloop.evaluate."""

    def __init__(self, store: FeatureStore, questions: QuestionSet):
        self.store, self.q, self.n = store, questions, questions.n

    def batch(self, s: int, e: int, device: torch.device) -> Dict[str, torch.Tensor]:
        rows = torch.tensor([self.store.index[v] for v in self.q.video[s:e].tolist()])
        a, v, p = self.store.layout.views(self.store.rows[rows])
        return dict(quest=self.q.tokens[s:e].to(device), audio=a.to(device), video=v.to(device), patch=p.to(device),
                    label=self.q.label[s:e].to(device), qtype_label=self.q.qtype[s:e].to(device))
