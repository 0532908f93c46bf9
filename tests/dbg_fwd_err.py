"""Debug helper (not a test): print full-forward logit errors per precision mode against the goldens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from oracle import synth
from oracle.make_golden import FWD_CASES
from gpu_util import build_model, dev, golden
for name in ("fwd_b8", "fwd_stress", "fwd_b256"):
    g = golden(name); B, T, E, K, ws, bs = FWD_CASES[name]
    sd = synth.make_state_dict(ws, num_experts=E); batch = synth.make_batch(bs, B, T)
    data = {k: v.to(dev()) for k, v in batch.items()}
    ref = torch.as_tensor(g["out"])
    for prec in ("fp32", "bf16"):
        m = build_model(sd, E, K, prec)
        for tt in ("fp32", "native", "tf32"):
            m.quest_encoder.compute = tt
            with torch.no_grad(): out = m(data)["out"].cpu()
            e = float((out - ref).norm() / ref.norm()); w = float(((out - ref).norm(dim=1) / ref.norm(dim=1)).max())
            print(f"{name} hot={prec} text={tt}: normwise {e:.2e} worst-sample {w:.2e} argmax-mismatch {(out.argmax(1) != ref.argmax(1)).sum().item()}")
