"""Stand-alone timing of b200_attention (CUDA events, L2-cold by rotating buffers): python tools/attn_bench.py [B N heads dh]."""
import sys

import torch

sys.path.insert(0, ".")
import b200path  # noqa: F401
import b200_native as nat

B, N, heads, dh = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 197, 12, 64)
C = heads * dh
nbuf = 4
qkv = [(torch.randn(B * N, 3 * C, device="cuda") * 0.9).bfloat16() for _ in range(nbuf)]
out = [torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
for i in range(nbuf):
    nat.attention(qkv[i], out[i], B, N, heads, dh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for i in range(reps):
    nat.attention(qkv[i % nbuf], out[i % nbuf], B, N, heads, dh)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 4.0 * B * heads * N * N * dh
bytes_ = B * N * 4 * C * 2
print(f"b200_attention B={B} N={N} heads={heads} dh={dh}: {ms * 1e3:.1f} us/launch, {flops / ms / 1e9:.1f} TFLOP/s, "
      f"{bytes_ / ms / 1e6:.0f} GB/s algorithmic, {ms * 1e3 / (B * heads):.2f} us per (case, head) over all SMs")
