"""Profiling aid (not a test): the tcgen05 GEMM kinds at the step's largest shapes, CUDA-graph timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qa_tiger_b200 import _cabi
dev = torch.device("cuda:0")
shapes = {"tf32_slf_in": (1, 57344, 1536, 512), "tf32_proj": (1, 114688, 512, 1024), "tf32_slf_out": (1, 57344, 512, 512),
          "x3_in_src": (2, 30720, 3584, 512), "x3_w1": (2, 15360, 1792, 512), "x3_w1_reduce": (2, 30720, 1792, 512),
          "tf32_proj_full": (1, 215040, 512, 1024), "f16_slf_in_full": (3, 215040, 1536, 512),
          "f16_slf_in": (3, 114688, 1536, 512), "f16_slf_out": (3, 114688, 512, 512), "f16_crs_kv": (3, 114688, 1024, 512)}
F16_OUT = {"tf32_proj", "f16_slf_in", "f16_slf_out", "f16_crs_kv", "tf32_proj_full", "f16_slf_in_full"}     # outputs that are fp16 planes on the product path
which = sys.argv[1:] or list(shapes)
for nm in which:
    kind, M, N, K = shapes[nm]
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5
    lin = _cabi.Lin(w, torch.randn(N, device=dev), _cabi.PREC_BF16, f16=True)
    planes = _cabi.split_planes(a) if kind == 2 else None
    if kind == 3: a = a.half()
    run = lambda: _cabi.gemm(kind, a, lin, a_planes=planes, out_f16=nm in F16_OUT)
    if nm == "x3_w1_reduce":           # expert first-layer GEMM with the Gaussian-weighted temporal reduction in its epilogue (EPI = 4)
        coef = torch.rand(M // 2, N // 256, device=dev)
        run = lambda: _cabi.gemm_reduce(a, lin, coef, 60, 256, a_planes=planes)
    for _ in range(2): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{nm}: [{M},{K}]x[{N},{K}]^T kind={kind}  {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s nominal")
