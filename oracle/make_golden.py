"""Golden-vector generator: runs the REAL reference (imported from /root/reference, read-only) on
seeded inputs and writes small fixtures to tests/golden/.  Run in the build container only:

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py fwd_b8     # one case

The reference cannot travel to the GPU box, the fixtures do.  Recipe for importing the reference
is SURVEY.md 8c: stub `ftfy`, replace `src.models.encoders.load` with a constructor of a
random-init CLIP with ViT-L/14@336px *text* dimensions, load tome/merge.py by file path.
Weights / inputs come from oracle/synth.py (seeded), so tests can regenerate the same inputs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import synth  # noqa: E402


def load_reference_model(E: int, K: int):
    sys.path.insert(0, REF)
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))
    import src.models.encoders as enc
    from src.models.base import clip_base

    enc.load = lambda name, q_aware_N=-1, device="cpu", **kw: (
        clip_base.CLIP(q_aware_N, 768, 28, 1, 64, 14, 77, 49408, 768, 12, 12).float().eval(), None)
    from src.models.net import QA_TIGER

    m = QA_TIGER(d_model=512, video_dim=768, patch_dim=1024, quest_dim=512, audio_dim=128, topK=K,
                 num_experts=E, encoder_type="ViT-L/14@336px").eval()
    return m


def load_reference_merge():
    spec = importlib.util.spec_from_file_location("ref_merge", os.path.join(REF, "src/tome/merge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sub(t: torch.Tensor, budget: int = 4096) -> np.ndarray:
    """Deterministic strided subsample of a block output (keeps fixtures small)."""
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // budget)
    return flat[::step].contiguous().numpy().copy()


FWD_CASES = {
    # name: (B, T, E, K, weight seed, batch seed)
    "fwd_b8": (8, 60, 7, 7, 713, 1),          # BASELINE config 1
    "fwd_b3": (3, 60, 7, 7, 713, 2),          # ragged last batch; quirk with B not dividing T
    "fwd_b1": (1, 60, 7, 7, 713, 3),          # quirk vanishes
    "fwd_k5": (5, 60, 10, 5, 714, 4),         # module defaults K<E
    "fwd_stress": (4, 120, 14, 14, 715, 5),   # BASELINE config 5 shape (T=120, E=K=14)
    "fwd_b256": (256, 60, 7, 7, 713, 1234),   # BASELINE config 2 batch (logits + ranks only)
    "fwd_stress_b128": (128, 120, 14, 14, 715, 6),   # BASELINE config 5 at its own per-GPU batch (logits + ranks only)
    "fwd_t150": (9, 150, 7, 7, 713, 7),       # more frames than one 128-token attention tile (logits + ranks only)
}


def gen_forward(name):
    B, T, E, K, wseed, bseed = FWD_CASES[name]
    sd = synth.make_state_dict(wseed, num_experts=E)
    m = load_reference_model(E, K)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    batch = synth.make_batch(bseed, B, T)
    cap = {}

    def hook(nm):
        def f(mod, inp, out):
            cap.setdefault(nm, []).append(out)
        return f

    for nm in ("crs_attn", "patch_selecter", "at_aggregator", "vt_aggregator", "quest_grounding",
               "quest_encoder"):
        getattr(m, nm).register_forward_hook(hook(nm))
    for agg in ("at_aggregator", "vt_aggregator"):
        mod = getattr(m, agg)
        orig_gg, orig_go = mod.generate_gaussian, mod.get_output

        def gg(pred, topk_inds, T=60, _o=orig_gg, _a=agg):
            w = _o(pred, topk_inds=topk_inds, T=T)
            cap[_a + ".topk_inds"] = topk_inds.clone()
            cap[_a + ".gauss_weight"] = w.clone()
            return w

        def go(el, gw, ti, tp, shape, _o=orig_go, _a=agg):
            cap[_a + ".topk_probs"] = tp.clone()
            return _o(el, gw, ti, tp, shape)

        mod.generate_gaussian, mod.get_output = gg, go
    with torch.no_grad():
        out = m({"quest": batch["quest"], "audio": batch["audio"], "video": batch["video"],
                 "patch": batch["patch"]})["out"]
    full = B <= 8
    d = dict(B=B, T=T, E=E, K=K, wseed=wseed, bseed=bseed, out=out.numpy(),
             argmax=out.argmax(1).numpy())
    for agg in ("at_aggregator", "vt_aggregator"):
        d[agg + ".topk_inds"] = cap[agg + ".topk_inds"].numpy().astype(np.int16)
        if full:
            d[agg + ".topk_probs"] = cap[agg + ".topk_probs"].numpy()
            d[agg + ".gauss_weight"] = cap[agg + ".gauss_weight"].numpy()
    if full:
        q, w = cap["quest_encoder"][0]
        d["quest768"], d["words768"] = sub(q), sub(w)
        a, v = cap["crs_attn"][0]
        d["audio_x"], d["video_x"] = sub(a), sub(v)
        ap, vp = cap["patch_selecter"][0]
        d["a_patch"], d["v_patch"] = sub(ap), sub(vp)
        d["a_global"] = cap["at_aggregator"][0].numpy()
        d["ap_global"], d["vp_global"] = (x.numpy() for x in cap["vt_aggregator"][0])
        d["fusion1"], d["fusion2"] = (x.numpy() for x in cap["quest_grounding"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "logits", out.shape, "argmax", out.argmax(1)[:8].tolist())


def gen_blocks():
    """Stand-alone block fixtures incl. the visualize=True weights and the standalone M2-M4."""
    sys.path.insert(0, REF)
    import src.models.modules as M
    sd = synth.make_state_dict(716, num_experts=7, text_tower=False)
    g = torch.Generator().manual_seed(6)
    B, T, D = 4, 60, 512
    audio, video = torch.randn(B, T, D, generator=g), torch.randn(B, T, D, generator=g)
    words = torch.randn(B, 77, D, generator=g)
    m = M.AVQCrossAttn(D, 8).eval()
    m.load_state_dict({k[len("crs_attn."):]: v for k, v in sd.items() if k.startswith("crs_attn.")})
    with torch.no_grad():
        a, v, (aw, vw) = m(audio, video, words, visualize=True)
    d = dict(avq_a=sub(a), avq_v=sub(v), avq_aw=sub(aw), avq_vw=sub(vw))
    # standalone get_output + generate_gaussian on random expert outputs (T,B,E,C)
    for tag, (B2, T2, E2, K2) in {"g7": (6, 60, 7, 7), "g10": (5, 60, 10, 5), "g14": (4, 120, 14, 14),
                                  "g1": (1, 60, 7, 7)}.items():
        moe = M.TempMoE(D, 8, topK=K2, n_experts=E2).eval()
        x = torch.randn(T2, B2, E2, D, generator=g)
        rl = torch.randn(B2, E2, generator=g)
        gr = torch.randn(B2, 2 * E2, generator=g)
        with torch.no_grad():
            probs = torch.softmax(rl, -1)
            tp, ti = torch.topk(probs, K2, dim=-1)
            tp = tp / tp.sum(-1, keepdim=True)
            gc = gr.view(B2, E2, 2).clone()
            gc[:, :, 0] = torch.tanh(gc[:, :, 0]) * moe.margin
            gc[:, :, 1] = torch.sigmoid(gc[:, :, 1])
            gw = moe.generate_gaussian(gc, topk_inds=ti, T=T2)
            o = moe.get_output(x, gw, ti, tp, (B2, T2, D))
        d.update({f"{tag}.out": o[:, 0].numpy(), f"{tag}.topk_inds": ti.numpy().astype(np.int16),
                  f"{tag}.topk_probs": tp.numpy(), f"{tag}.gauss_weight": gw.numpy(),
                  f"{tag}.shape": np.array([B2, T2, E2, K2])})
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **d)
    print("blocks ok")


def gen_tome():
    rm = load_reference_merge()
    d = {}
    # (1) full chain 577 -> 14 on 2 frames, stand-in metric, merge_wavg + merge_source
    g = torch.Generator().manual_seed(42)
    N, t0, C = 2, 577, 1024
    x = torch.randn(N, t0, C, generator=g)
    size, source = None, None
    sched = [25] * 23 + [0]
    k = 0
    for r in sched:
        metric = x.view(N, x.shape[1], C // 64, 64).mean(2)
        merge, _ = rm.bipartite_soft_matching(metric, r, class_token=True, distill_token=False)
        if merge is rm.do_nothing:
            continue
        # recover the indices the closure captured
        cells = {n: c.cell_contents for n, c in zip(merge.__code__.co_freevars, merge.__closure__)}
        d[f"chain.{k}.unm"] = cells["unm_idx"][..., 0].numpy().astype(np.int16)
        d[f"chain.{k}.src"] = cells["src_idx"][..., 0].numpy().astype(np.int16)
        d[f"chain.{k}.dst"] = cells["dst_idx"][..., 0].numpy().astype(np.int16)
        source = rm.merge_source(merge, x, source)
        x, size = rm.merge_wavg(merge, x, size)
        k += 1
    d["chain.steps"] = np.array(k)
    d["chain.x"] = x.numpy()
    d["chain.size"] = size.numpy()
    d["chain.group"] = source.argmax(dim=1).numpy().astype(np.int16)     # [N,577] -> group position
    assert x.shape == (N, 14, C)
    # (2) single steps on odd shapes / no class token / tiny t
    cases = {"s_a": (3, 27, 64, 25, True), "s_b": (2, 52, 64, 25, True), "s_c": (2, 10, 64, 3, False),
             "s_d": (1, 3, 64, 25, True), "s_e": (2, 577, 64, 25, True), "s_f": (2, 9, 64, 100, False)}
    for nm, (n, t, c, r, cls) in cases.items():
        metric = torch.randn(n, t, c, generator=g)
        xx = torch.randn(n, t, 32, generator=g)
        merge, _ = rm.bipartite_soft_matching(metric, r, class_token=cls, distill_token=False)
        cells = {k2: c2.cell_contents for k2, c2 in zip(merge.__code__.co_freevars, merge.__closure__)}
        xo, so = rm.merge_wavg(merge, xx, None)
        d[f"{nm}.cfg"] = np.array([n, t, c, r, int(cls)])
        d[f"{nm}.metric"], d[f"{nm}.x"] = metric.numpy(), xx.numpy()
        for kk in ("unm", "src", "dst"):
            d[f"{nm}.{kk}"] = cells[kk + "_idx"][..., 0].numpy().astype(np.int16)
        d[f"{nm}.xo"], d[f"{nm}.so"] = xo.numpy(), so.numpy()
    np.savez_compressed(os.path.join(OUT, "tome.npz"), **d)
    print("tome ok", k, "steps")
    # (3) distillation token (merge.py:56-57, 77-78; never set by the extractor): single steps, with and without the class token
    dd = {}
    cases = {"d_a": (3, 27, 64, 25, True), "d_b": (2, 52, 64, 10, False), "d_c": (2, 577, 64, 25, True), "d_d": (1, 4, 64, 25, True),
             "d_e": (2, 197, 64, 25, True)}
    for nm, (n, t, c, r, cls) in cases.items():
        metric = torch.randn(n, t, c, generator=g)
        xx = torch.randn(n, t, 32, generator=g)
        merge, _ = rm.bipartite_soft_matching(metric, r, class_token=cls, distill_token=True)
        cells = {k2: c2.cell_contents for k2, c2 in zip(merge.__code__.co_freevars, merge.__closure__)}
        xo, so = rm.merge_wavg(merge, xx, None)
        dd[f"{nm}.cfg"] = np.array([n, t, c, r, int(cls)])
        dd[f"{nm}.metric"], dd[f"{nm}.x"] = metric.numpy(), xx.numpy()
        for kk in ("unm", "src", "dst"):
            dd[f"{nm}.{kk}"] = cells[kk + "_idx"][..., 0].numpy().astype(np.int16)
        dd[f"{nm}.xo"], dd[f"{nm}.so"] = xo.numpy(), so.numpy()
        dd[f"{nm}.mean"] = merge(xx, mode="mean").numpy()
        dd[f"{nm}.group"] = rm.merge_source(merge, xx, None).argmax(dim=1).numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "tome_distill.npz"), **dd)
    print("tome distill ok")


def load_reference_tome_patch():
    """The REAL tome/patch/timm.py, imported by file path with `timm.models.vision_transformer` stubbed (oracle/timm_stub.py:
    timm is not installed here) and `src.tome` registered as a bare package (its __init__ pulls in PIL-based visualisation)."""
    from oracle import timm_stub
    timm_stub.install()
    sys.path.insert(0, REF)
    if "src.tome" not in sys.modules:
        src = sys.modules.get("src") or types.ModuleType("src")
        src.__path__ = [os.path.join(REF, "src")]
        sys.modules["src"] = src
        t = types.ModuleType("src.tome")
        t.__path__ = [os.path.join(REF, "src", "tome")]
        sys.modules["src.tome"] = t
    spec = importlib.util.spec_from_file_location("ref_tome_timm", os.path.join(REF, "src/tome/patch/timm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, timm_stub


def gen_tome_block():
    """two ToMe-patched ViT-L blocks (dim 1024, 16 heads) of the real reference on 2 frames of 197 tokens, r = 25
    each (the second block therefore sees size != 1: proportional attention), trace_source on; plus a 577-token single block."""
    mod, stub = load_reference_tome_patch()
    sd = synth.make_vit_blocks(717, 2)
    d = {}
    for nm, (N, t, rs) in {"b197": (2, 197, [25, 25]), "b577": (1, 577, [25])}.items():
        blocks = []
        info = {"r": list(rs), "size": None, "source": None, "trace_source": True, "prop_attn": True, "class_token": True, "distill_token": False}
        for i in range(len(rs)):
            b = stub.Block(1024, 16)
            b.load_state_dict({k[len(f"blocks.{i}."):]: v for k, v in sd.items() if k.startswith(f"blocks.{i}.")})
            b.__class__ = mod.ToMeBlock
            b.attn.__class__ = mod.ToMeAttention
            b._tome_info = info
            blocks.append(b.eval())
        g = torch.Generator().manual_seed(1000 + t)
        x = torch.randn(N, t, 1024, generator=g)
        d[f"{nm}.cfg"] = np.array([N, t, len(rs)] + list(rs))
        with torch.no_grad():
            for i, b in enumerate(blocks):
                xa, metric = b.attn(b.norm1(x), info["size"])
                d[f"{nm}.{i}.attn"], d[f"{nm}.{i}.metric"] = sub(xa), metric.numpy()
                x = b(x)
                d[f"{nm}.{i}.x"] = sub(x, 8192)
                d[f"{nm}.{i}.size"] = info["size"].numpy()
        d[f"{nm}.x_final"] = sub(x, 65536)
        d[f"{nm}.group"] = info["source"].argmax(dim=1).numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "tome_block.npz"), **d)
    print("tome_block", {k: v.shape for k, v in d.items() if k.endswith("x_final")})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or list(FWD_CASES) + ["blocks", "tome", "tome_block"]
    torch.set_num_threads(os.cpu_count())
    for w in which:
        if w in FWD_CASES:
            gen_forward(w)
        elif w == "blocks":
            gen_blocks()
        elif w == "tome":
            gen_tome()
        elif w == "tome_block":
            gen_tome_block()
        else:
            raise SystemExit("unknown case " + w)
