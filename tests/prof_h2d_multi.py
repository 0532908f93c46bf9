"""Profiling aid (not a test): concurrent pinned host -> device copies on N ranks (torchrun), with and without binding each
rank to its GPU's local CPUs / NUMA node -- what bounds bench.py's e2e leg at N > 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from qa_tiger_b200 import hostbind
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 881 * (1 << 20) // 4
for mode in ("unbound", "bound"):
    if mode == "bound":
        prev = hostbind.bind_to_gpu(lr)
        print(f"rank {rank}: local cpus {hostbind.gpu_local_cpus(lr)[:4] if hostbind.gpu_local_cpus(lr) else None}.. prev affinity size {len(prev) if prev else None}", flush=True)
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 8], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{mode}: {world} ranks x 881 MB concurrently: {ms.item():.2f} ms per copy (max over ranks) = {world * 0.881 / (ms.item() * 1e-3) / 1e0:.1f} GB/s aggregate", flush=True)
    del h, d
dist.destroy_process_group()
