"""Profiling aid (not a test): hot-path throughput with one CUDA graph on one stream against two graphs (two independent
batches, each with its own workspace) replayed on two streams -- how much of the step's ramps, tails and latency-bound
launches another batch in flight can fill."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from qa_tiger_b200.models.net import QA_TIGER

dev = torch.device("cuda:0")
B, T, P, E = int(os.environ.get("B", "256")), 60, 14, 7
K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
sd = synth.make_state_dict(713, num_experts=E)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, audio_dim=128, topK=E, num_experts=E)
model.load_state_dict(sd)
model = model.to(dev).eval().set_precision("bf16")
batches = []
for i in range(2):
    g = torch.Generator(device=dev).manual_seed(1234 + i)
    audio, video = torch.randn(B, T, 128, device=dev, generator=g), torch.randn(B, T, 768, device=dev, generator=g)
    patch = torch.randn(B, T, P, 1024, device=dev, generator=g)
    quest = synth.make_batch(1234 + i, B, T, with_patch=False)["quest"].to(dev)
    with torch.no_grad():
        q768, w768 = model.quest_encoder(quest)
    batches.append((q768.contiguous(), w768.contiguous(), audio, video, patch))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
graphs, outs = [], []
with torch.no_grad():
    for i in range(2):
        with torch.cuda.stream(streams[i]):
            for _ in range(3):
                ref = model.forward_features(*batches[i])
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=streams[i]):
            o = model.forward_features(*batches[i])
        graphs.append(gr); outs.append((o, ref))
torch.cuda.synchronize()


def run(two, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = torch.cuda.Event()
    torch.cuda.synchronize()
    a.record(streams[0])
    streams[1].wait_event(a)
    for i in range(n):
        k = i % 2 if two else 0
        with torch.cuda.stream(streams[k]):
            graphs[i % 2].replay() if two else graphs[0].replay()
    done.record(streams[1])
    streams[0].wait_event(done)
    b.record(streams[0])
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for two in (False, True, False, True):
    run(two, 6)
    ms = run(two, K)
    print(f"{'two streams' if two else 'one stream '}: {ms:.3f} ms per step of {B} questions = {B / ms:.1f} k questions/s")
for o, ref in outs:
    print("bit-identical to the eager result:", bool(torch.equal(o, ref)))
