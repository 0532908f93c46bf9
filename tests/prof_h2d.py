"""Profiling aid (not a test): pinned host -> device copy bandwidth (what bounds bench.py's e2e leg)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (64, 256, 881):
    n = mb * (1 << 20) // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    h.normal_()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"H2D pinned {mb} MB: {ms:.2f} ms  {mb / 1024 / (ms * 1e-3):.1f} GiB/s", flush=True)
    # two halves on two streams
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
        with torch.cuda.stream(s2): d[n // 2:].copy_(h[n // 2:], non_blocking=True)
    torch.cuda.synchronize()
    ms2 = (time.perf_counter() - t0) / 5 * 1e3
    print(f"   2 streams: {ms2:.2f} ms  {mb / 1024 / (ms2 * 1e-3):.1f} GiB/s", flush=True)
    d2 = torch.empty(1 << 20, device=dev)
    hh = torch.empty(1 << 20).pin_memory()
    e0.record()
    for _ in range(5): hh.copy_(d2, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"   D2H 4 MB: {e0.elapsed_time(e1) / 5:.3f} ms")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv", shell=True, capture_output=True, text=True).stdout)
print(subprocess.run("lscpu | head -20; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)
