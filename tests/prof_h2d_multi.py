"""Profiling aid (not a test): concurrent pinned host -> device copies on N ranks (torchrun) -- what bounds the e2e legs at N > 1.
Variants: torch's pinned allocator (cudaHostAlloc default flags), write-combined pinned memory (cudaHostAllocWriteCombined),
a 2 MiB-aligned malloc registered with cudaHostRegister (THP-eligible), four concurrent chunked copies on four streams, and a
subset of the ranks copying alone (does one GPU's rate survive its neighbours?)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cudart = ctypes.CDLL("libcudart.so")
NBYTES = 881 << 20


def host_wc(nbytes):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))      # cudaHostAllocWriteCombined
    assert rc == 0, rc
    return p


def host_registered(nbytes):
    libc = ctypes.CDLL("libc.so.6")
    libc.aligned_alloc.restype = ctypes.c_void_p
    p = ctypes.c_void_p(libc.aligned_alloc(ctypes.c_size_t(1 << 21), ctypes.c_size_t(nbytes)))
    libc.madvise(p, ctypes.c_size_t(nbytes), 14)                                                     # MADV_HUGEPAGE
    ctypes.memset(p, 1, nbytes)
    rc = cudart.cudaHostRegister(p, ctypes.c_size_t(nbytes), ctypes.c_uint(0))
    assert rc == 0, rc
    return p


def timed(copy_fn, label, active=True):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(6): copy_fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 6 if active else 0.0], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_active = torch.tensor([1.0 if active else 0.0], device="cuda")
    dist.all_reduce(n_active)
    if rank == 0:
        print(f"{label}: {int(n_active.item())} ranks x 881 MB: {ms.item():.2f} ms per copy (max) = {n_active.item() * NBYTES / 1e9 / (ms.item() * 1e-3):.1f} GB/s aggregate", flush=True)


d = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
h = torch.empty(NBYTES, dtype=torch.uint8).pin_memory(); h.fill_(1)
timed(lambda: d.copy_(h, non_blocking=True), "torch pinned")
streams = [torch.cuda.Stream() for _ in range(4)]
q = NBYTES // 4
def chunked():
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            d[i * q:(i + 1) * q].copy_(h[i * q:(i + 1) * q], non_blocking=True)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
timed(chunked, "torch pinned, 4 streams x 220 MB")
for k in (1, 2, 4):
    if k < world:
        timed(lambda: d.copy_(h, non_blocking=True), f"torch pinned, only ranks < {k} copy", active=rank < k)
        timed(lambda: d.copy_(h, non_blocking=True), f"torch pinned, only every {world // k}-th rank copies", active=rank % (world // k) == 0)
for name, mk in (("write-combined cudaHostAlloc", host_wc), ("cudaHostRegister of a THP-advised malloc", host_registered)):
    p = mk(NBYTES)
    stream = torch.cuda.current_stream().cuda_stream
    fn = lambda: cudart.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(NBYTES), ctypes.c_int(1), ctypes.c_void_p(stream))
    timed(fn, name)
if rank == 0:
    os.system("nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)' ; grep -i hugepages /proc/meminfo | head -4; free -g | head -2")
dist.destroy_process_group()
