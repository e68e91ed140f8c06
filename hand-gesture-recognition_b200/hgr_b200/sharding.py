"""Batch sharding for multi-GPU inference.

Every hand crop is independent (eval-mode BatchNorm uses running statistics,
attention stays inside one image), so the batch dimension is split into
contiguous shards, one per rank, with no data-path collective.  The only
cross-rank traffic is the barrier and the max-over-ranks reduction of the timed
region that bench.py needs.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device_node(device_index: int) -> int | None:
    """One process per GPU: run this process (and, by first touch, the pinned staging buffers it allocates afterwards)
    on the CPU socket the GPU hangs off, so that the per-step host->device copies of eight ranks do not all cross the
    socket interconnect.  Reads the GPU's NUMA node from sysfs and narrows the process's CPU affinity to that node's
    cores (never widens it).  Returns the node, or None when the platform does not say (single-socket boxes report
    -1) - then nothing is changed."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id  # torch >= 2.6
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # no such attribute / sysfs entry / device: leave the affinity alone
        return None


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of `rank`'s contiguous shard; the first `total % world` ranks hold one extra unit."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (elapsed milliseconds); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
