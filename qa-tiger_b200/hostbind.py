"""Host-side placement for the input pipe: bind a rank's CPU threads -- and with them its first-touch /
pinned allocations -- to the NUMA node its GPU hangs off, so the per-step host -> device copy does not cross the socket
interconnect.  NVML names the CPUs local to a device; the process affinity is narrowed to them before the pinned staging
buffers are allocated.  A no-op (returns None) when NVML or the affinity call is unavailable or the box has one node."""
from __future__ import annotations

import os
from typing import List, Optional


def _nvml_handle(device_index: int):
    import pynvml
    import torch

    pynvml.nvmlInit()
    props = torch.cuda.get_device_properties(device_index)
    uuid = getattr(props, "uuid", None)
    if uuid is not None:
        try:
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:  # noqa: BLE001 -- fall back to the PCI address
            pass
    bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
    return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs NVML reports as local to the device (None when it cannot tell)."""
    try:
        pynvml, h = _nvml_handle(device_index)
        n = os.cpu_count() or 1
        words = (n + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [i for i in range(n) if (int(mask[i // 64]) >> (i % 64)) & 1]
        return cpus or None
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never a requirement
        return None


def bind_to_gpu(device_index: int) -> Optional[List[int]]:
    """Narrow this process's CPU affinity to the device's local CPUs.  Returns the previous affinity (to restore with
    os.sched_setaffinity(0, prev)) or None when nothing was changed."""
    if os.environ.get("QAT_HOSTBIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return None
    prev = sorted(os.sched_getaffinity(0))
    want = sorted(set(cpus) & set(prev))
    if not want or want == prev:
        return None
    os.sched_setaffinity(0, want)
    return prev
