"""Data-parallel eval driver: the B200 counterpart of /root/reference/src/trainutils.py:397-462 (`test`).

What changes against the reference loop:
  * one process per GPU, contiguous shards [r*ceil(N/W), (r+1)*ceil(N/W)) of the question list, local batches
    walked in order, no shuffle and no wrap-around padding (the reference's DistributedSampler(shuffle=True)
    double-counts padded samples, trainutils.py:192).  Because logits depend on the LOCAL batch,
    this partition is part of the parity contract: `local_batches()` is what an oracle must be evaluated on.
  * accuracy counters live on the device as one packed int64[20] = (correct, total, tot[9], correct[9]);
    no per-sample `.item()` (trainutils.py:419-423 does B+1 host syncs per batch).
  * collectives, once at the end: all_gather of the int64 predictions (padded to the shard length) and ONE
    all_reduce(SUM) of the packed counters, replacing the 20 scalar all-reduces of trainutils.py:436-440.
The forward itself never communicates.
"""
from __future__ import annotations

import logging
import math
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch
import torch.distributed as dist

# trainutils.py / dataset.py:22-31 question-type table (modality -> type -> counter slot)
QTYPE2IDX = {
    "Audio": {"Counting": 0, "Comparative": 1},
    "Visual": {"Counting": 2, "Location": 3},
    "Audio-Visual": {"Existential": 4, "Counting": 5, "Location": 6, "Comparative": 7, "Temporal": 8},
}
N_QTYPES = 9


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    per = math.ceil(n / world) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def local_batches(n: int, rank: int, world: int, batch_size: int) -> List[Tuple[int, int]]:
    """[start, end) global index ranges of this rank's local batches (last one ragged)."""
    lo, hi = shard_range(n, rank, world)
    return [(s, min(hi, s + batch_size)) for s in range(lo, hi, batch_size)]


class SyntheticAVQA:
    """MUSIC-AVQA-test-sized synthetic set: sample i is a pure function of (seed, i), so
    every sharding sees identical samples.  Features are generated on `device` batch by batch."""

    def __init__(self, n: int = 9129, seed: int = 1234, T: int = 60, P: int = 14, n_unique_questions: int = 0):
        self.n, self.seed, self.T, self.P = n, seed, T, P
        self.n_unique = n_unique_questions     # >0: draw questions from a pool (MUSIC-AVQA test has 33 unique strings)

    def _tokens(self, i: int) -> torch.Tensor:
        key = i if self.n_unique <= 0 else (i * 2654435761) % self.n_unique
        g = torch.Generator().manual_seed(self.seed * 1000003 + key)
        q = torch.zeros(77, dtype=torch.int32)
        L = int(torch.randint(5, 20, (1,), generator=g))
        q[0] = 49406
        q[1:1 + L] = torch.randint(1, 49406, (L,), generator=g, dtype=torch.int32)
        q[1 + L] = 49407
        return q

    def batch(self, start: int, end: int, device: torch.device) -> Dict[str, torch.Tensor]:
        B = end - start
        audio = torch.empty(B, self.T, 128, device=device)
        video = torch.empty(B, self.T, 768, device=device)
        patch = torch.empty(B, self.T, self.P, 1024, device=device)
        g = torch.Generator(device=device)
        lab = torch.empty(B, dtype=torch.int64)
        qt = torch.empty(B, dtype=torch.int64)
        for j, i in enumerate(range(start, end)):
            g.manual_seed(self.seed * 7919 + i)
            audio[j].normal_(generator=g)
            video[j].normal_(generator=g)
            patch[j].normal_(generator=g)
            h = (i * 1103515245 + self.seed) & 0x7FFFFFFF
            lab[j], qt[j] = h % 42, (h // 42) % N_QTYPES
        quest = torch.stack([self._tokens(i) for i in range(start, end)])
        return dict(quest=quest.to(device), audio=audio, video=video, patch=patch, label=lab.to(device), qtype_label=qt.to(device))


def pack_counters(pred: torch.Tensor, target: torch.Tensor, qtype: torch.Tensor, acc: torch.Tensor) -> None:
    """acc int64[20] += (correct, total, tot[9], correct[9]) -- all on `acc.device`, no host sync."""
    hit = (pred == target).to(torch.int64)
    acc[0] += hit.sum()
    acc[1] += pred.numel()
    acc[2:2 + N_QTYPES].index_add_(0, qtype, torch.ones_like(hit))
    acc[2 + N_QTYPES:].index_add_(0, qtype, hit)


@torch.no_grad()
def evaluate(forward_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], dataset, device: torch.device, batch_size: int = 256,
             rank: int = 0, world: int = 1, group=None, logger: Optional[logging.Logger] = None) -> Dict[str, object]:
    """forward_fn(batch dict) -> logits [B,42].  Returns accuracy, the packed counters and ALL predictions (global order)."""
    n = dataset.n
    lo, hi = shard_range(n, rank, world)
    per = math.ceil(n / world)
    counters = torch.zeros(2 + 2 * N_QTYPES, dtype=torch.int64, device=device)
    preds = torch.full((per,), -1, dtype=torch.int64, device=device)
    # a dataset with `iter_batches` (datapipe.StreamingLoader) prefetches one batch ahead on its own streams; a plain one
    # is asked for each local batch in turn
    if hasattr(dataset, "iter_batches"):
        batches = dataset.iter_batches(rank, world, batch_size, device)
    else:
        batches = ((s, e, dataset.batch(s, e, device)) for (s, e) in local_batches(n, rank, world, batch_size))
    for s, e, b in batches:
        logits = forward_fn(b)
        p = logits.argmax(dim=1)
        preds[s - lo:e - lo] = p
        pack_counters(p, b["label"].reshape(-1), b["qtype_label"].reshape(-1), counters)
    if world > 1:
        gathered = torch.empty(world * per, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(gathered, preds, group=group)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        all_preds = torch.cat([gathered[r * per:r * per + (shard_range(n, r, world)[1] - shard_range(n, r, world)[0])] for r in range(world)])
    else:
        all_preds = preds[:hi - lo]
    c = counters.cpu()                                                 # the single host sync of the eval
    correct, total = int(c[0]), int(c[1])
    result = dict(acc=100.0 * correct / max(total, 1), correct=correct, total=total, tot=c[2:2 + N_QTYPES].tolist(),
                  correct_by_type=c[2 + N_QTYPES:].tolist(), predictions=all_preds.cpu())
    if logger is not None and rank == 0:
        for line in format_report(result):
            logger.info(line)
    return result


def format_report(result: Dict[str, object]) -> List[str]:
    """Same line formats as trainutils.py:442-461 so result files stay comparable."""
    lines = []
    tot, cor = result["tot"], result["correct_by_type"]
    for modality, types in QTYPE2IDX.items():
        mc = mt = 0
        for qst_type, idx in types.items():
            c, t = cor[idx], tot[idx]
            mc, mt = mc + c, mt + t
            value = c / t * 100.0 if t else float("nan")
            key = f"{modality}/{qst_type}"
            lines.append(f"Test {key:>24} accuracy: {value:.2f}({c}/{t})")
        macc = mc / mt * 100.0 if mt else float("nan")
        lines.append(f"Test {modality:>24} accuracy: {macc:.2f}({mc}/{mt})")
    key = "Total avg"
    lines.append(f"Test {key:>24} accuracy: {result['acc']:.2f}({result['correct']}/{result['total']})")
    return lines


# ---------------------------------------------------------------------------------------------------------------
# `python -m qa_tiger_b200.evalloop` (or under torchrun, one rank per GPU): BASELINE.json configs[3], the data-parallel
# eval over a synthetic MUSIC-AVQA-test-sized set.  The shard is generated on the device up front (resident inputs:
# 9 129 questions x 3.65 MB = 33 GB fit one B200), then the eval loop is timed: text tower + hot path + argmax +
# on-device counters per local batch, one all_gather + one all_reduce at the end.
def _main():
    import argparse
    import json
    import os
    import time
    import warnings

    from .models.net import QA_TIGER

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=9129)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--unique-questions", type=int, default=33, help="question pool size (MUSIC-AVQA test: 33 unique strings); 0 = all distinct")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--text", default="native", choices=["native", "fp32", "bf16x3"])
    ap.add_argument("--no-dedup", action="store_true")
    ap.add_argument("--no-memo", action="store_true", help="do not remember encoded questions across batches")
    ap.add_argument("--stream", default="", choices=["", "fp16", "fp32"],
                    help="evaluate from a PINNED HOST per-video feature store through datapipe.StreamingLoader (device cache, prefetch, "
                         "tower on a side stream) instead of device-resident batches; the value is the patch dtype of the store")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("qa_tiger_b200.evalloop needs a GPU (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from . import hostbind
    hostbind.bind_to_gpu(local)          # host threads / pinned buffers on the GPU's NUMA node (no-op on one node)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # the version banner goes to stdout regardless of the file
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(713)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, audio_dim=128, topK=7, num_experts=7)
    model = model.to(dev).eval().set_precision(args.precision)
    model.quest_encoder.compute = args.text
    model.quest_encoder.dedup = not args.no_dedup
    model.quest_encoder.memo = not (args.no_memo or args.no_dedup)
    fwd = lambda b: model(b)["out"]
    if args.stream:
        from . import datapipe as dp
        qs = dp.QuestionSet.synthetic_music_avqa(n=args.n, n_unique_questions=args.unique_questions)
        lo, hi = shard_range(qs.n, rank, world)
        t0 = time.perf_counter()
        store = dp.FeatureStore.synthetic(qs.videos_of(lo, hi), dp.RowLayout(patch_dtype=torch.float16 if args.stream == "fp16" else torch.float32),
                                          device=dev)
        torch.cuda.synchronize()
        t_gen = time.perf_counter() - t0
        Resident = dp.StreamingLoader(store, qs, dev, args.batch, rank, world, encoder=model.quest_encoder)
        cache = local_batches(args.n, rank, world, args.batch)
    else:
        ds = SyntheticAVQA(args.n, n_unique_questions=args.unique_questions)
        t0 = time.perf_counter()
        cache = {(s, e): ds.batch(s, e, dev) for (s, e) in local_batches(args.n, rank, world, args.batch)}
        torch.cuda.synchronize()
        t_gen = time.perf_counter() - t0

        class Resident:
            n = args.n

            @staticmethod
            def batch(s, e, device):
                return cache[(s, e)]

    evaluate(fwd, Resident, dev, args.batch, rank, world)          # warm-up pass (weight preparation, workspaces)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = evaluate(fwd, Resident, dev, args.batch, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"workload": f"DP eval over {args.n} synthetic questions (configs[3]), per-GPU batch {args.batch}, "
                                      f"question pool {args.unique_questions or 'all distinct'}, text tower {args.text}"
                                      f"{'' if args.no_dedup else ' + dedup'}{' + cross-batch memo' if model.quest_encoder.memo else ''}",
                          "input": (f"pinned host store ({args.stream} patches), datapipe.StreamingLoader, second pass (every video resident "
                                    "in the device cache)") if args.stream else "device-resident batches",
                          "n_gpus": world, "questions": args.n, "wall_s": dt, "questions_per_s": args.n / dt,
                          "local_batches_rank0": len(cache), "generate_s_rank0": t_gen, "acc_vs_random_labels": res["acc"],
                          "memo_entries_rank0": len(getattr(model.quest_encoder, "_memo_state", {}).get("slots", {})) if model.quest_encoder.memo else 0,
                          "note": "timed pass follows one warm-up pass over the same shard (weights prepared, workspaces sized, question memo filled)",
                          "total": res["total"], "collectives": "1 all_gather_into_tensor + 1 all_reduce(int64[20])" if world > 1 else "none"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    _main()
