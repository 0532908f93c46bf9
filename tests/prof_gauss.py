"""Profiling aid (not a test): the fused Gaussian-expert kernel at config-2 size, fp32 and bf16 inputs.
Launches are replayed from a CUDA graph so the CPU-side ctypes/torch overhead (~30 us per call) is not timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qa_tiger_b200.models.modules import gauss_reduce
dev = torch.device("cuda:0")
B, T, E, K, C = (int(x) for x in os.environ.get("GAUSS_SHAPE", "256,60,7,7,512").split(","))
ts = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rl, gr = torch.randn(B, E, device=dev), torch.randn(B, 2 * E, device=dev)
for dt in (torch.float32, torch.bfloat16):
    xs = [torch.randn(T, B, E, C, device=dev).to(dt) for _ in range(3)]   # rotate 3 buffers (> L2 in total)
    for i in range(3):
        gauss_reduce(xs[i % 3], rl, gr, K, t_split=ts)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(9):
            gauss_reduce(xs[i % 3], rl, gr, K, t_split=ts)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 45
    byts = xs[0].numel() * xs[0].element_size()
    print(f"{dt} B={B} T={T} E={E} t_split={ts}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s  ({byts/ms/1e6/6536.4*100:.1f}% of measured HBM peak)")
