"""Sharding of the image batch across ranks (one process per GPU, no data-path collective).

The path shards by image: rank r of R owns a contiguous block of the batch.  The
only collective traffic is control-plane: a barrier around the timed region and a max-reduce of
the per-rank times.  Backend-agnostic (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the items rank `rank` owns; blocks differ by at most one item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(item: int, n_items: int, world: int) -> int:
    base, rem = divmod(n_items, world)
    edge = rem * (base + 1)
    if item < edge:
        return item // (base + 1)
    return rem + (item - edge) // base if base else world - 1


class Timing:
    """Barrier + max-over-ranks helpers around an (optional) torch.distributed process group."""

    def __init__(self, dist=None, device=None):
        self.dist, self.device = dist, device

    @property
    def world(self) -> int:
        return self.dist.get_world_size() if self.dist is not None else 1

    @property
    def rank(self) -> int:
        return self.dist.get_rank() if self.dist is not None else 0

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def aggregate_throughput(units_per_rank: float, seconds_this_rank: float, timing: Timing) -> float:
    """Whole-job throughput = units all ranks processed / max-over-ranks time."""
    total = timing.sum_over_ranks(units_per_rank)
    t = timing.max_over_ranks(seconds_this_rank)
    return total / t if t > 0 else 0.0


# ---- host placement: staging buffers of rank r belong on the NUMA node its GPU hangs off ---------------
def _parse_cpulist(s: str):
    cpus = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_node(pci_bus_id: str) -> int:
    """NUMA node of a GPU from sysfs (-1 when the platform does not say)."""
    import os
    bus = pci_bus_id.lower()
    if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:  # nvidia-smi style 00000000:1B:00.0
        bus = bus[4:]
    try:
        with open(os.path.join("/sys/bus/pci/devices", bus, "numa_node")) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def bind_host_to_gpu(device_index: int) -> dict:
    """Prefer the GPU's NUMA node for this process's CPU time and (pinned) host allocations.

    An 8-GPU box feeds every GPU over its own PCIe root port; when all ranks stage their covers in one
    socket's memory, the DMA of the GPUs on the other socket crosses the inter-socket link and the whole
    job saturates there.  Best effort: returns what was done and never raises."""
    import ctypes
    import os
    info = {"numa_node": -1, "cpus": 0, "mempolicy": False}
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = gpu_numa_node(bus)
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            want = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
        # set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode): x86_64 syscall 238
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(16 * 64))
        info["mempolicy"] = rc == 0
    except Exception:
        pass
    return info
