"""CPU: the C-ABI library loads and exports every symbol include/qat.h declares; the Python structs
have the C layout; host-side guards fail loudly without a GPU (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qat.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|size_t|const char\*|long long)\s+(qat_\w+)\s*\(", src, re.M)))


def test_header_symbols_are_exported():
    from qa_tiger_b200 import _cabi
    import qa_tiger_b200.build as b
    b.build()
    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
        assert n in _cabi.EXPORTS, f"{n} declared in qat.h but not bound in _cabi.py"
    assert lib.qat_abi_version() == _cabi.ABI_VERSION == 4
    assert lib.qat_tome_clamp_r(577, 25, 1) == 25 and lib.qat_tome_clamp_r(27, 25, 1) == 13 and lib.qat_tome_clamp_r(1, 25, 1) == 0


def test_struct_layout_matches_c(tmp_path):
    """Compile a tiny C program against qat.h that prints sizeof/offsetof and compare with ctypes."""
    from qa_tiger_b200 import _cabi
    pairs = {"QatLinearW": _cabi.LinearW, "QatGemmArgs": _cabi.GemmArgs, "QatLinearArgs": _cabi.LinearArgs, "QatAvqArgs": _cabi.AvqArgs,
             "QatPatchArgs": _cabi.PatchArgs, "QatGaussArgs": _cabi.GaussArgs, "QatMoeArgs": _cabi.MoeArgs,
             "QatGroundArgs": _cabi.GroundArgs, "QatTomeMatchArgs": _cabi.TomeMatchArgs,
             "QatTomeMergeArgs": _cabi.TomeMergeArgs, "QatTomeChainArgs": _cabi.TomeChainArgs,
             "QatClipLayerW": _cabi.ClipLayerW, "QatClipTextArgs": _cabi.ClipTextArgs, "QatAttnArgs": _cabi.AttnArgs, "QatGemmAttnArgs": _cabi.GemmAttnArgs,
             "QatForwardArgs": _cabi.ForwardArgs, "QatGatherArgs": _cabi.GatherArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qat.h"', 'int main(void){']
    for cname, ct in pairs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in ct._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in pairs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for f, _ in ct._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(ct, f).offset, f"{cname}.{f}"


def test_options_are_set_through_the_abi_not_the_environment(monkeypatch):
    """A/B switches: explicit setter, snapshot per call; a leaked environment variable after load changes nothing, and the
    product library has no GEMM debug mode (QAT_GEMM_DBG is compiled out: no symbol, no getenv on the hot path)."""
    from qa_tiger_b200 import _cabi
    lib = _cabi.lib()
    assert _cabi.get_option("gemm_cg") == 0
    monkeypatch.setenv("QAT_GEMM_CG", "1")
    assert _cabi.get_option("gemm_cg") == 0
    with _cabi.option("gemm_cg", 2):
        assert _cabi.get_option("gemm_cg") == 2
    assert _cabi.get_option("gemm_cg") == 0
    with pytest.raises(RuntimeError, match="unknown option"):
        _cabi.set_option("gemm_dbg", 1)
    blob = open(_cabi.LIB_PATH, "rb").read()
    assert b"QAT_GEMM_DBG" not in blob
    assert lib.qat_abi_version() == _cabi.ABI_VERSION == 4


def test_workspace_sizing_runs_without_gpu():
    from qa_tiger_b200 import _cabi
    lib = _cabi.lib()
    a = _cabi.AvqArgs()
    a.B, a.T, a.L, a.D, a.H, a.prec = 256, 60, 77, 512, 8, 0
    for f, n, k in (("in_src", 3584, 512), ("in_words", 1024, 512), ("out_cat", 512, 1536), ("ffn1", 512, 512), ("ffn2", 512, 512)):
        w = getattr(a, f)
        w.w, w.n, w.k = 256, n, k          # dummy non-NULL pointer: sizing never dereferences it
    a.norm1.gamma = a.norm1.beta = a.norm2.gamma = a.norm2.beta = 256
    need = lib.qat_workspace_bytes(_cabi.OP_AVQ, C.addressof(a))
    assert need > 2 * 256 * 60 * 3584 * 4
    a.D = 500
    assert lib.qat_workspace_bytes(_cabi.OP_AVQ, C.addressof(a)) == 0
    assert b"H*64" in lib.qat_last_error()


def test_cpu_tensors_are_rejected_loudly():
    from qa_tiger_b200.models import modules as M
    m = M.Projection(128, 512).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(4, 128))
    moe = M.TempMoE(512, 8, topK=7, n_experts=7).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        moe(torch.randn(2, 512), torch.randn(2, 60, 512))
    with pytest.raises(RuntimeError, match="inference-only"):
        M.AVQCrossAttn(512, 8).train()(torch.randn(1, 60, 512), torch.randn(1, 60, 512), torch.randn(1, 77, 512))


def test_missing_library_is_an_error(monkeypatch):
    from qa_tiger_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libqat_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "qa-tiger_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "/root/reference" not in txt.replace(
                    "/root/reference/src", "REFDOC"), f
