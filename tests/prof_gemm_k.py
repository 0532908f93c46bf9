"""Profiling aid (not a test): per-tile cost model of the tcgen05 GEMM -- sweep K at fixed M, N and fit
cycles/tile = a + b*K; QAT_GEMM_DBG switches isolate epilogue / MMA / store costs
(they exist only in a profile build of the library: NVCC_EXTRA="-DQAT_PROFILE_BUILD" python qa-tiger_b200/build.py --force)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qa_tiger_b200 import _cabi
dev = torch.device("cuda:0")
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 3
M, N = int(os.environ.get("M", 114688)), int(os.environ.get("N", 1536))
out_f16 = os.environ.get("OUT", "f16") == "f16"
KS = [int(k) for k in os.environ.get("KS", "64,128,256,512,1024,2048").split(",")]
for K in KS:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5
    lin = _cabi.Lin(w, torch.randn(N, device=dev), _cabi.PREC_BF16, f16=True)
    planes = _cabi.split_planes(a) if kind == 2 else None
    ah = a.half() if kind == 3 else a
    run = lambda: _cabi.gemm(kind, ah, lin, a_planes=planes, out_f16=out_f16)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    per_sm = -(-tiles // 148)
    print(f"kind={kind} K={K:5d}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.0f} TFLOP/s  {ms*1e3/per_sm:6.2f} us/tile ({per_sm} tiles/SM)", flush=True)
