"""Profiling aid (not a test): host-side cost of one QA_TIGER.forward (ctypes marshalling, weight-cache checks,
workspace sizing) -- what bounds the eval loop once the text tower is deduplicated."""
import cProfile, pstats, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from qa_tiger_b200.models.net import QA_TIGER
dev = torch.device("cuda:0")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, audio_dim=128, topK=7, num_experts=7).to(dev).eval().set_precision("bf16")
m.quest_encoder.compute = "native"; m.quest_encoder.dedup = True
b = {k: v.to(dev) for k, v in synth.make_batch(1, 256, 60).items()}
with torch.no_grad():
    for _ in range(3): m(b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): m(b)
    t_issue = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    print(f"cpu issue per forward: {t_issue*1e3:.2f} ms")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): m(b)
    pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
