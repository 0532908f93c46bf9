"""Debug helper (not a test): run one block per process argument and print errors."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import qat_oracle as O, synth
from gpu_util import dev, relerr, sub_sd
from qa_tiger_b200.models import modules as M
which = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
sd = synth.make_state_dict(716, num_experts=7, text_tower=False)
g = torch.Generator().manual_seed(0)
B, T = 2, 60
if which == "avq":
    m = M.AVQCrossAttn(512, 8); m.load_state_dict(sub_sd(sd, "crs_attn")); m = m.to(dev()).eval(); m.precision = prec
    a, v, w = torch.randn(B, T, 512, generator=g), torch.randn(B, T, 512, generator=g), torch.randn(B, 77, 512, generator=g)
    ra, rv = O.avq_cross_attn(sd, "crs_attn", a, v, w)
    ga, gv = m(a.to(dev()), v.to(dev()), w.to(dev())); torch.cuda.synchronize()
    print("avq", relerr(ga, ra), relerr(gv, rv))
elif which == "lin":
    m = M.Projection(768, 512); m.load_state_dict(sub_sd(sd, "video_proj")); m = m.to(dev()).eval(); m.precision = prec
    x = torch.randn(4096, 768, generator=g)
    ref = O.linear(x, sd["video_proj.proj.weight"], sd["video_proj.proj.bias"])
    got = m(x.to(dev())); torch.cuda.synchronize()
    print("lin", relerr(got, ref))
elif which == "gauss":
    x = torch.randn(60, 6, 7, 512, generator=g); rl = torch.randn(6, 7, generator=g); gr = torch.randn(6, 14, generator=g)
    ref, aux = O.gauss_reduce(x, rl, gr, 7)
    for ts in (1, 4):
        out = M.gauss_reduce(x.to(dev()), rl.to(dev()), gr.to(dev()), 7, t_split=ts); torch.cuda.synchronize()
        print("gauss tsplit", ts, relerr(out, ref))
elif which == "tome":
    from qa_tiger_b200.tome import merge as tm
    from oracle import tome_oracle as TO
    metric = torch.randn(2, 27, 64, generator=g); x = torch.randn(2, 27, 32, generator=g)
    res = TO.match(metric, 25, True)
    idx = tm.match(metric.to(dev()), 25, True); torch.cuda.synchronize()
    print("tome", [a.cpu().tolist() == b.tolist() for a, b in zip(idx, res[:3])])
