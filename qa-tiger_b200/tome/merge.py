"""Drop-in mirror of /root/reference/src/tome/merge.py on the B200 path.

  bipartite_soft_matching(metric, r, class_token, distill_token) -> (merge, unmerge)   merge.py:18-97
  merge_wavg(merge, x, size) -> (x, size)                                              merge.py:210-225
  merge_source(merge, x, source) -> source                                             merge.py:228-241
plus the index-form API the kernels really speak (`match`, `merge_with`, `chain`).

Indices are bit-identical to the reference's (tests/test_gpu_tome.py); they are int64 `[N,k,1]` tensors
on the closures (`merge.unm_idx`, `.src_idx`, `.dst_idx`) exactly like the reference's closure cells.
`distill_token=True` (merge.py:56-57, 77-78; the extractor never sets it, timm.py:48-53) is carried as bit 1 of the C ABI's
protected-token flag word; `unmerge` keeps the reference's own (distill-unaware) row split.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch

from .. import _cabi as cabi

MODES = {"wavg": 0, "sum": 1, "mean": 2, "amax": 3}


def do_nothing(x, mode=None):
    return x


def _flags(class_token, distill_token=False) -> int:
    return int(bool(class_token)) | (2 if distill_token else 0)


def clamp_r(t: int, r: int, class_token: bool = False, distill_token: bool = False) -> int:
    return cabi.lib().qat_tome_clamp_r(int(t), int(r), _flags(class_token, distill_token))


def match(metric: torch.Tensor, r: int, class_token: bool = False, distill_token: bool = False):
    """metric [N,t,64] fp32 -> (unm_idx [N,ta-r], src_idx [N,r], dst_idx [N,r]) int32, or None when r clamps to 0."""
    cabi.require_cuda_f32(metric, "metric")
    N, t, Cc = metric.shape
    rr = clamp_r(t, r, class_token, distill_token)
    if rr <= 0:
        return None
    ta = (t + 1) // 2
    dev = metric.device
    unm = torch.empty(N, ta - rr, dtype=torch.int32, device=dev)
    src = torch.empty(N, rr, dtype=torch.int32, device=dev)
    dst = torch.empty(N, rr, dtype=torch.int32, device=dev)
    a = cabi.TomeMatchArgs()
    a.N, a.t, a.C, a.r, a.class_token = N, t, Cc, r, _flags(class_token, distill_token)
    a.stream, a.metric = cabi.stream_ptr(dev), metric.data_ptr()
    a.unm_idx, a.src_idx, a.dst_idx = unm.data_ptr(), src.data_ptr(), dst.data_ptr()
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().qat_tome_match(C.byref(a)), "qat_tome_match")
    return unm, src, dst


def merge_with(x: torch.Tensor, idx, mode: str = "wavg", size: Optional[torch.Tensor] = None,
               group: Optional[torch.Tensor] = None, distill_token: bool = False):
    """Apply one merge step.  mode 'wavg' returns (x_out, size_out [N,t-r,1]); other modes return x_out."""
    unm, src, dst = idx
    cabi.require_cuda_f32(x, "x")
    N, t, Cc = x.shape
    r = src.shape[1]
    dev = x.device
    out = torch.empty(N, t - r, Cc, dtype=torch.float32, device=dev)
    a = cabi.TomeMergeArgs()
    a.N, a.t, a.C, a.r, a.mode, a.stream = N, t, Cc, r, MODES[mode], cabi.stream_ptr(dev)
    a.distill_token = int(bool(distill_token))
    a.x, a.unm_idx, a.src_idx, a.dst_idx, a.x_out = x.data_ptr(), unm.data_ptr(), src.data_ptr(), dst.data_ptr(), out.data_ptr()
    size_out = None
    if mode == "wavg":
        if size is not None:
            size = size.reshape(N, t).contiguous()
            cabi.require_cuda_f32(size, "size", (N, t))
            a.size = size.data_ptr()
        size_out = torch.empty(N, t - r, dtype=torch.float32, device=dev)
        a.size_out = size_out.data_ptr()
    if group is not None:
        if group.dtype != torch.int32 or not group.is_contiguous():
            raise RuntimeError("group must be a contiguous int32 [N,T0] tensor")
        a.group, a.T0 = group.data_ptr(), group.shape[1]
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().qat_tome_merge(C.byref(a)), "qat_tome_merge")
    return (out, size_out[..., None]) if mode == "wavg" else out


def bipartite_soft_matching(metric: torch.Tensor, r: int, class_token: bool = False,
                            distill_token: bool = False) -> Tuple[Callable, Callable]:
    idx = match(metric, r, class_token, distill_token)
    if idx is None:
        return do_nothing, do_nothing
    unm, src, dst = idx
    t = metric.shape[1]
    rr = src.shape[1]

    def merge(x: torch.Tensor, mode="mean") -> torch.Tensor:
        return merge_with(x.contiguous().float(), idx, mode, distill_token=distill_token)

    def unmerge(x: torch.Tensor) -> torch.Tensor:                     # merge.py:82-95 (not on the hot path)
        unm_len = unm.shape[1]
        u, d = x[..., :unm_len, :], x[..., unm_len:, :]
        n, _, c = u.shape
        s = d.gather(dim=-2, index=dst.long()[..., None].expand(n, rr, c))
        out = torch.zeros(n, t, c, device=x.device, dtype=x.dtype)
        out[..., 1::2, :] = d
        out.scatter_(dim=-2, index=(2 * unm.long())[..., None].expand(n, unm_len, c), src=u)
        out.scatter_(dim=-2, index=(2 * src.long())[..., None].expand(n, rr, c), src=s)
        return out

    merge.idx = idx
    merge.distill_token = bool(distill_token)
    merge.unm_idx, merge.src_idx, merge.dst_idx = (i.long()[..., None] for i in idx)
    return merge, unmerge


def merge_wavg(merge: Callable, x: torch.Tensor, size: Optional[torch.Tensor] = None):
    if merge is do_nothing:
        if size is None:
            size = torch.ones_like(x[..., 0, None])
        return x, size
    return merge_with(x.contiguous().float(), merge.idx, "wavg", size, distill_token=getattr(merge, "distill_token", False))


def merge_source(merge: Callable, x: torch.Tensor, source: Optional[torch.Tensor] = None) -> torch.Tensor:
    if source is None:
        n, t, _ = x.shape
        source = torch.eye(t, device=x.device)[None, ...].expand(n, t, t)
    if merge is do_nothing:
        return source
    return merge(source.contiguous(), mode="amax")


def default_schedule(steps: int = 23, r: int = 25, blocks: int = 24) -> List[int]:
    """extract_tome14.py:162 (`model.r = [25]*23`) padded with zeros to the 24 ViT-L blocks (tome/utils.py:94-97)."""
    return [r] * steps + [0] * (blocks - steps)


def chain(x: torch.Tensor, schedule: Optional[List[int]] = None, class_token: bool = True, trace: bool = False,
          want_group: bool = True):
    """Run the whole merge schedule (T0) with the stand-in metric of SURVEY 8d.
    Returns (x_out [N,t_end,C], size [N,t_end,1], group [N,t0] int32 | None, steps | None) where steps is a list of
    (unm_idx, src_idx, dst_idx) int32 tensors per executed step."""
    cabi.require_cuda_f32(x, "x")
    N, t0, Cc = x.shape
    if schedule is None:
        schedule = default_schedule()
    l = cabi.lib()
    n = len(schedule)
    rl = (C.c_int32 * max(n, 1))(*schedule)
    t_end = l.qat_tome_chain_t_end(t0, n, rl, int(class_token))
    dev = x.device
    out = torch.empty(N, t_end, Cc, dtype=torch.float32, device=dev)
    size = torch.empty(N, t_end, dtype=torch.float32, device=dev)
    group = torch.empty(N, t0, dtype=torch.int32, device=dev) if want_group else None
    tr = None
    if trace:
        tr = torch.empty(max(1, l.qat_tome_chain_trace_len(N, t0, n, rl, int(class_token))), dtype=torch.int32, device=dev)
    a = cabi.TomeChainArgs()
    a.N, a.t0, a.C, a.n_steps, a.r_list, a.class_token = N, t0, Cc, n, rl, int(class_token)
    a.x, a.x_out, a.size_out, a.group, a.idx_trace = x.data_ptr(), out.data_ptr(), size.data_ptr(), cabi.ptr(group), cabi.ptr(tr)
    cabi.run(cabi.OP_TOME_CHAIN, "qat_tome_chain", a, dev)
    steps = None
    if trace:
        steps, off, t = [], 0, t0
        for r in schedule:
            rr = clamp_r(t, r, class_token)
            if rr <= 0:
                continue
            n_unm = (t + 1) // 2 - rr
            unm = tr[off:off + N * n_unm].view(N, n_unm); off += N * n_unm
            src = tr[off:off + N * rr].view(N, rr); off += N * rr
            dst = tr[off:off + N * rr].view(N, rr); off += N * rr
            steps.append((unm, src, dst))
            t -= rr
    return out, size[..., None], group, steps


def frame_shard(n_frames: int, rank: int, world: int):
    """contiguous shard [lo, hi) of the frame list owned by `rank` (ToMe shards by frame, no exchange)"""
    per = -(-n_frames // max(world, 1))
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


_side_streams = {}


def _streams_for(dev: torch.device, n: int):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _side_streams.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def chain_frames(frames, n_frames: int, rank: int = 0, world: int = 1, chunk: int = 592, schedule: Optional[List[int]] = None,
                 class_token: bool = True, want_group: bool = False, streams: int = 2):
    """The extractor's loop (scripts/extract_ToMe/extract_tome14.py:153-170: one frame at a time through the 23-step schedule)
    over this rank's shard of `n_frames` frames, `chunk` frames per `chain()` call (a multiple of the SM count keeps the
    one-frame-per-CTA match kernel in whole waves).  `frames`: a [n_frames, t0, C] CUDA tensor, or a callable
    (lo, hi) -> [hi-lo, t0, C] producing the frames of a chunk (x0 of configs[2] is 36 GB: callers may stream it).
    `streams` > 1: consecutive chunks run on that many side streams (each with its own workspace), so the FFMA-bound match
    kernels of one chunk overlap the HBM-bound merge kernels of another; the chunks are independent, results are unchanged.
    Returns (x_out [n_local, t_end, C], size [n_local, t_end, 1], group [n_local, t0] | None) in frame order."""
    lo, hi = frame_shard(n_frames, rank, world)
    outs, sizes, groups = [], [], []
    starts = list(range(lo, hi, chunk))
    dev = frames.device if isinstance(frames, torch.Tensor) else torch.device("cuda", torch.cuda.current_device())
    side = _streams_for(dev, streams) if streams > 1 and len(starts) > 1 else None
    if side is not None:
        cur = torch.cuda.current_stream(dev)
        for st in side:
            st.wait_stream(cur)                     # the frames (and anything else queued so far) are ready
    for i, c0 in enumerate(starts):
        c1 = min(hi, c0 + chunk)
        if side is None:
            x = frames(c0, c1) if callable(frames) else frames[c0:c1]
            o, sz, g, _ = chain(x, schedule, class_token, trace=False, want_group=want_group)
        else:
            with torch.cuda.stream(side[i % len(side)]):
                x = frames(c0, c1) if callable(frames) else frames[c0:c1]
                o, sz, g, _ = chain(x, schedule, class_token, trace=False, want_group=want_group)
        outs.append(o)
        sizes.append(sz)
        groups.append(g)
    if side is not None:
        for st in side:
            cur.wait_stream(st)                     # outputs are consumed on the caller's stream
        for t_ in outs + sizes + [g for g in groups if g is not None]:
            t_.record_stream(cur)
    if not outs:
        return None, None, None
    return torch.cat(outs), torch.cat(sizes), (torch.cat(groups) if want_group else None)
