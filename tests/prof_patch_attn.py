"""Times the PatchSelecter block (patch_proj + self-attention + cross-attention kernels) at the bench shape with the A-resident
fused kernels (default) and with the shared-memory-operand form (option gemm_attn_no_ts = 1).  python tests/prof_patch_attn.py [B T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from oracle import synth
from gpu_util import dev, relerr, sub_sd
from qa_tiger_b200 import _cabi
from qa_tiger_b200.models import modules as M


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    sd = synth.make_state_dict(716, num_experts=7, text_tower=False)
    m = M.PatchSelecter(512, 8)
    m.load_state_dict(sub_sd(sd, "patch_selecter"))
    m = m.to(dev()).eval()
    m.precision = "bf16"
    proj = M.Projection(1024, 512)
    proj.load_state_dict(sub_sd(sd, "patch_proj"))
    proj = proj.to(dev()).eval()
    proj.precision = "bf16"
    g = torch.Generator().manual_seed(1)
    raw = torch.randn(B, T, 14, 1024, generator=g).to(dev())
    a, v = torch.randn(B, T, 512, generator=g).to(dev()), torch.randn(B, T, 512, generator=g).to(dev())
    outs = {}
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for name, val in (("a_resident", 0), ("smem_operands", 1), ("a_resident", 0), ("smem_operands", 1)):
        with _cabi.option("gemm_attn_no_ts", val), _cabi.option("gemm_attn_ts_flags", flags), torch.no_grad():
            for _ in range(3):
                o = m(raw, a, v, proj=proj)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                o = m(raw, a, v, proj=proj)
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:14s} block {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us", flush=True)
            outs[name] = o
    print("equal:", torch.equal(outs["a_resident"][0], outs["smem_operands"][0]), torch.equal(outs["a_resident"][1], outs["smem_operands"][1]),
          relerr(outs["a_resident"][0], outs["smem_operands"][0]), relerr(outs["a_resident"][1], outs["smem_operands"][1]))


if __name__ == "__main__":
    main()
