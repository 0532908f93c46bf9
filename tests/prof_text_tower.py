"""Profiling aid (not a test): the native CLIP text tower on 256 distinct questions (qat_clip_text_fwd), eager launches."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from qa_tiger_b200.models.net import QA_TIGER

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = synth.make_state_dict(713, num_experts=7)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, audio_dim=128, topK=7, num_experts=7)
model.load_state_dict(sd)
model = model.to(dev).eval().set_precision("bf16")
quest = synth.make_batch(1234, B, 60, with_patch=False)["quest"].to(dev)
with torch.no_grad():
    for _ in range(2):
        model.quest_encoder(quest)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        model.quest_encoder(quest)
    b.record(); torch.cuda.synchronize()
print(f"text tower B={B}: {a.elapsed_time(b) / 5:.3f} ms per call")
