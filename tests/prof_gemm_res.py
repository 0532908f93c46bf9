import os, sys
sys.path.insert(0, "/root/repo")
import torch
from qa_tiger_b200 import _cabi
dev = torch.device("cuda:0")
M, N, K = 114688, 512, 512
a = torch.randn(M, K, device=dev).half(); w = torch.randn(N, K, device=dev) * K ** -0.5
lin = _cabi.Lin(w, torch.randn(N, device=dev), _cabi.PREC_BF16, f16=True)
res = torch.randn(M, N, device=dev).half()
for label, kw in (("no res", {}), ("res fp16", {"res": res})):
    run = lambda: _cabi.gemm(3, a, lin, out_f16=True, **kw)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{label}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
