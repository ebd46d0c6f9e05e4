"""Host-to-device bandwidth of pinned buffers allocated under each candidate CPU set (sharding._bind_by_probe)."""
import sys
sys.path.insert(0, ".")
import b200path  # noqa: F401
import torch
from sharding import bind_to_device_numa_node, restore_affinity
import os
print("allowed cpus:", sorted(os.sched_getaffinity(0)))
info = bind_to_device_numa_node(0)
print(info)
n = 256 << 20
host = torch.empty(n, dtype=torch.uint8, pin_memory=True); host.fill_(1)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
dst.copy_(host, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): dst.copy_(host, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D after binding: %.1f GB/s" % (8 * n / (e0.elapsed_time(e1) * 1e6)))
restore_affinity(info)
