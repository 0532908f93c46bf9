"""Profiling aid (not a test): ToMe 577->14 chain (config 3) throughput on the GPU, frames/s, vs the CPU oracle port."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qa_tiger_b200.tome import merge as tm
from qa_tiger_b200 import _cabi
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 960
x = torch.randn(N, 577, 1024, device=dev, generator=torch.Generator(device="cuda").manual_seed(42))
for _ in range(2):
    out = tm.chain(x, want_group=True)
torch.cuda.synchronize()
_cabi.lib().qat_launch_count(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 3
for _ in range(reps):
    out = tm.chain(x, want_group=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
launches = _cabi.lib().qat_launch_count(0) // reps
step_bytes = 56.49e6 * N       # stepwise merge traffic per frame
print(f"tome chain N={N}: {ms:.2f} ms  {N/ms*1e3:.0f} frames/s  launches={launches}  stepwise-traffic roofline {step_bytes/ms/1e6:.0f} GB/s "
      f"({step_bytes/ms/1e6/6536.4*100:.1f}% of measured HBM peak)")
if "--cpu" in sys.argv:
    from oracle import tome_oracle as TO
    xc = x[:16].cpu()
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter(); TO.chain(xc); dt = time.perf_counter() - t0
    print(f"cpu oracle port: 16 frames in {dt:.2f} s = {16/dt:.1f} frames/s on {torch.get_num_threads()} threads")
