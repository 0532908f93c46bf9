import os
import warnings

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def dev():
    return torch.device("cuda:0")


def build_model(sd, E, K, precision="fp32"):
    from qa_tiger_b200.models.net import QA_TIGER
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, quest_dim=512, audio_dim=128, topK=K, num_experts=E)
    m.load_state_dict(sd)
    return m.to(dev()).eval().set_precision(precision)


def sub_sd(sd, prefix):
    return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


# gates (BASELINE.json north_star): logits within 1e-5 (fp32 mode) / 1e-3 (bf16 mode), normwise relative
TOL = {"fp32": 1e-5, "bf16": 1e-3}
