"""Debug aid (not a test): the whole forward against the oracle over unusual shapes (tiny / odd T, P, B, E, K)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import qat_oracle as O
from oracle import synth
from gpu_util import build_model, dev, relerr

cases = [(1, 1, 14, 7, 7), (1, 2, 14, 7, 7), (2, 3, 1, 7, 3), (3, 5, 2, 3, 1), (5, 17, 16, 5, 5), (2, 33, 7, 9, 4), (7, 60, 14, 1, 1),
         (2, 61, 14, 7, 7), (1, 128, 3, 7, 7), (2, 129, 5, 4, 2), (300, 8, 2, 7, 7), (4, 31, 14, 14, 14)]
for prec in ("fp32", "bf16"):
    for (B, T, P, E, K) in cases:
        try:
            sd = synth.make_state_dict(713, num_experts=E)
            model = build_model(sd, E, K, prec)
            batch = synth.make_batch(B * 7 + T, B, T, P=P)
            with torch.no_grad():
                ref = O.qa_tiger_forward(sd, batch, K=K, E=E)
                out = model({k: v.to(dev()) for k, v in batch.items()})["out"]
            torch.cuda.synchronize()
            e = relerr(out, ref)
            same = out.argmax(1).cpu().tolist() == ref.argmax(1).tolist()
            print(f"{prec} B={B} T={T} P={P} E={E} K={K}: relerr {e:.2e} argmax {'same' if same else 'DIFF'}", flush=True)
        except Exception as ex:   # noqa: BLE001
            print(f"{prec} B={B} T={T} P={P} E={E} K={K}: ERROR {str(ex)[:300]}", flush=True)
