"""Stand-alone timing of b200_scale_map at the C3 shapes (B = 1024, 32 x 32 maps): python tools/scale_bench.py"""
import sys
sys.path.insert(0, ".")
import b200path  # noqa: F401
import torch
import b200_native as nat
B = 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for C in (128, 256, 512):
    x = torch.randn(B, 32, 32, C, device="cuda").bfloat16()
    y = torch.empty_like(x)
    gate = torch.rand(B, C, device="cuda")
    nat.scale_map(x, y, gate=gate)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nat.scale_map(x, y, gate=gate); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 5
    print(f"scale_map C={C}: {ms:.3f} ms  {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
