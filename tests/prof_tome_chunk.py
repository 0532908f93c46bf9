"""Profiling aid (not a test): ToMe 577->14 chain throughput against the chunk size and the number of chunk streams
(measured: 0.57 / 0.63 of the stepwise-traffic roofline at 592 frames per chunk on one / two streams, 0.61 / 0.63 at 2 368)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qa_tiger_b200.tome import merge as tm
dev = torch.device("cuda:0")
N = 4736
x = torch.empty(N, 577, 1024, device=dev)
g = torch.Generator(device=dev).manual_seed(42)
for c0 in range(0, N, 592):
    x[c0:c0 + 592].normal_(generator=g)
for chunk in (592, 1184, 2368, 296, 592):
    for streams in (1, 2):
        tm.chain_frames(x, N, chunk=chunk, streams=streams)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            tm.chain_frames(x, N, chunk=chunk, streams=streams)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 2
        print(f"chunk={chunk} streams={streams}: {N / ms:.1f} k frames/s  roofline frac {56.49e6 * N / ms / 1e6 / 6536.4:.3f}", flush=True)
