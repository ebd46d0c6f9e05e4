"""Case sharding across GPUs.

Every case is independent at inference (BatchNorm uses running statistics; nothing in the path
mixes cases), so a batch is cut into contiguous per-rank shards, weights are replicated and the
only exchange is one all_gather of the [B/n, K] fp32 logits.  The reference has no distributed
code; this is the B200 arrangement of its single-device loop.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_cases: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`; the first n % world ranks take one extra case."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_cases, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_logits(local_logits: torch.Tensor, n_cases: int):
    """All ranks call this with their shard's logits; returns the [n_cases, K] logits in case order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    sizes = [shard_bounds(n_cases, r, world)[1] - shard_bounds(n_cases, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = local_logits.new_zeros((pad, local_logits.shape[1]))
    buf[: local_logits.shape[0]] = local_logits
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def max_over_ranks(value: float, device) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa_node(device_index: int):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned host buffers
    it allocates afterwards - and the thread that issues the copies - sit next to the PCIe root of that GPU.  On a
    two-socket host a buffer on the far socket caps host-to-device copies at the inter-socket link (measured on this
    pool: 10.7 GB/s instead of > 18 GB/s, which makes the end-to-end path copy bound).  Returns a dict describing
    what was done; never raises (containers often hide the topology: then nothing is changed)."""
    import os

    info = {"node": None, "cpus": None, "bound": False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return _bind_by_probe(device_index, info)
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        info["cpus"] = len(target)
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
            info["previous"] = allowed
    except Exception as exc:  # topology not visible: leave the affinity alone
        info["error"] = type(exc).__name__
    return info


def _bind_by_probe(device_index: int, info):
    """sysfs does not say which node the GPU hangs off (containers report -1): measure it.  For every visible NUMA
    node (or, failing that, each half of the allowed CPUs) pin the process there, allocate a pinned buffer (its pages
    land on the local node), time a host-to-device copy, and stay on the fastest candidate if it beats the slowest by
    more than 15 %."""
    import glob
    import os

    allowed = os.sched_getaffinity(0)
    cands = []
    for path in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        try:
            with open(path) as f:
                c = _parse_cpulist(f.read()) & allowed
            if c:
                cands.append(c)
        except Exception:
            pass
    if len(cands) < 2:
        ordered = sorted(allowed)
        if len(ordered) < 4:
            return info
        cands = [set(ordered[:len(ordered) // 2]), set(ordered[len(ordered) // 2:])]
    dev = torch.device("cuda", device_index)
    rates = []
    try:
        rates = [0.0] * len(cands)
        for rep_i in range(2 * len(cands)):  # every candidate twice, interleaved; best of the two (the first copies of a
            i = rep_i % len(cands)           # process also pay one-off driver work)
            c = cands[i]
            os.sched_setaffinity(0, c)
            n = (48 + 4 * rep_i) << 20  # distinct sizes: the caching host allocator must not hand back another probe's block
            host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            host.fill_(1)
            dst = torch.empty(n, dtype=torch.uint8, device=dev)
            dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                dst.copy_(host, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            rates[i] = max(rates[i], 4 * n / (e0.elapsed_time(e1) * 1e6))
            del host, dst
        info["probe_gbps"] = [round(r, 1) for r in rates]
        best = max(range(len(cands)), key=lambda i: rates[i])
        if rates[best] > 1.15 * min(rates):
            os.sched_setaffinity(0, cands[best])
            info.update(cpus=len(cands[best]), bound=True, previous=allowed, node=f"probe:{best}")
        else:
            os.sched_setaffinity(0, allowed)
    except Exception as exc:
        info["error"] = type(exc).__name__
        try:
            os.sched_setaffinity(0, allowed)
        except Exception:
            pass
    return info


def restore_affinity(info):
    import os

    if info.get("bound") and info.get("previous"):
        try:
            os.sched_setaffinity(0, info["previous"])
        except Exception:
            pass
