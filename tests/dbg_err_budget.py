"""Debug helper (not a test): precision-policy study on the CPU oracle.

Emulates operand / storage formats of the CUDA path by rounding the operands (and optionally the
stored outputs) of tagged contractions inside oracle.qat_oracle, then prints the logit error against
the fp32 oracle.  Usage: python tests/dbg_err_budget.py [B]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import qat_oracle as O
from oracle import synth


def rn_fp16(x): return x.half().float()
def rn_bf16(x): return x.bfloat16().float()
def tf32_trunc(x): return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def tf32_rn(x):
    i = x.view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
def bf16x2(x):
    hi = x.bfloat16().float()
    return hi + (x - hi).bfloat16().float()
def ident(x): return x


PATCH_TAGS = ("patch_proj", "patch_selecter.slf", "patch_selecter.crs.k", "patch_selecter.crs.v")


def is_patch(tag):
    return tag.startswith(PATCH_TAGS)


def run(policy, B=16, T=60, E=7):
    """policy(tag) -> (fx, fw, fout) rounding functions for operand x, weight w and the stored output."""
    sd = synth.make_state_dict(713, num_experts=E)
    batch = synth.make_batch(1234, B, T)
    base_linear = O.linear.__wrapped__ if hasattr(O.linear, "__wrapped__") else O.linear

    def linear(x, w, b, tag=""):
        fx, fw, fo = policy(tag)
        y = fx(x.contiguous()) @ fw(w.contiguous()).t()
        if b is not None:
            y = y + b
        return fo(y.contiguous())

    linear.__wrapped__ = base_linear
    O.linear = linear
    try:
        with torch.no_grad():
            out = O.qa_tiger_forward(sd, batch, K=E, E=E)
    finally:
        O.linear = base_linear
    return out


def report(name, out, ref):
    e = float((out - ref).norm() / ref.norm())
    w = float(((out - ref).norm(dim=1) / ref.norm(dim=1)).max())
    mm = int((out.argmax(1) != ref.argmax(1)).sum())
    print(f"{name:70s} normwise {e:.2e}  worst-sample {w:.2e}  argmax-mismatch {mm}", flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.set_num_threads(os.cpu_count())
    fp32 = lambda tag: (ident, ident, ident)
    ref = run(fp32, B)

    def chain(tag):            # non-patch contractions as today: bf16x3 ~ bf16 hi+lo operands
        if tag.startswith("fp32") or tag.startswith("text"):
            return (ident, ident, ident)
        return (bf16x2, bf16x2, ident)

    def mk(patch_fmt, store=ident, store_tags=()):
        def pol(tag):
            if is_patch(tag):
                fo = store if tag.startswith(store_tags) else ident
                return (patch_fmt, patch_fmt, fo)
            return chain(tag)
        return pol

    report("serial bf16x3 + patch tf32 (truncating, = round-1 hardware path)", run(mk(tf32_trunc), B), ref)
    report("serial bf16x3 + patch tf32 RN", run(mk(tf32_rn), B), ref)
    report("serial bf16x3 + patch fp16 RN operands, fp32 storage", run(mk(rn_fp16), B), ref)
    report("  + qkv/kv stored fp16", run(mk(rn_fp16, rn_fp16, ("patch_selecter.slf.q", "patch_selecter.slf.k", "patch_selecter.slf.v", "patch_selecter.crs.k", "patch_selecter.crs.v")), B), ref)
    report("  + everything on the patch path stored fp16 (incl. p, p')", run(mk(rn_fp16, rn_fp16, PATCH_TAGS), B), ref)
    report("serial bf16x3 + patch bf16 RN operands", run(mk(rn_bf16), B), ref)
    # patch_proj in tf32 straight from the fp32 input, rest fp16
    def mixed(tag):
        if tag.startswith("patch_proj"):
            return (tf32_trunc, tf32_trunc, ident)
        return mk(rn_fp16, rn_fp16, ("patch_selecter.slf.q", "patch_selecter.slf.k", "patch_selecter.slf.v", "patch_selecter.crs.k", "patch_selecter.crs.v"))(tag)
    report("patch_proj tf32-trunc, selecter fp16 operands + fp16 qkv/kv storage", run(mixed, B), ref)
    def allfp16(tag):
        if tag.startswith("fp32") or tag.startswith("text"):
            return (ident, ident, ident)
        return (rn_fp16, rn_fp16, ident)
    report("everything fp16 RN single pass (router chain fp32)", run(allfp16, B), ref)

    # ---- round-1b: which other contractions tolerate single-pass fp16 (RN) operands? ----
    P16 = PATCH_TAGS + ("patch_selecter.crs.q", "patch_selecter.crs.o", "patch_selecter.mlp0", "patch_selecter.mlp2")
    def pol_sel(extra=(), store=()):
        def pol(tag):
            if tag.startswith("patch_proj"):
                return (tf32_trunc, tf32_trunc, rn_fp16)
            if tag.startswith(P16) or tag.startswith(extra):
                return (rn_fp16, rn_fp16, rn_fp16 if tag.startswith(PATCH_TAGS + store) else ident)
            return chain(tag)
        return pol
    report("current GPU path (proj tf32, slf/kv fp16 planes)", run(mk_cur := (lambda tag: (tf32_trunc, tf32_trunc, rn_fp16) if tag.startswith("patch_proj") else ((rn_fp16, rn_fp16, rn_fp16) if is_patch(tag) else chain(tag))), B), ref)
    report("+ whole PatchSelecter fp16 single pass (crs.q/o, mlp0/2)", run(pol_sel(), B), ref)
    report("+ experts first layer (e1) fp16 single pass", run(pol_sel((("at_aggregator.e1", "vt_aggregator.e1"))), B), ref)
    report("+ experts e1 only (selecter small GEMMs stay bf16x3)", run(lambda tag: (rn_fp16, rn_fp16, ident) if tag.endswith(".e1") else mk_cur(tag), B), ref)
