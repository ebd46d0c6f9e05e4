"""Run one GEMM shape of the path a few times (target for `ncu -k regex:conv_gemm -s N -c 1`).

    python tools/one_gemm.py M K N [gelu|f32res]
    python tools/one_gemm.py conv B H W Cin Cout taps [gap]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200path  # noqa: F401,E402
import b200_native as nat  # noqa: E402

if sys.argv[1] == "conv":
    B, H, W, CI, CO, TAPS = (int(a) for a in sys.argv[2:8])
    x = (torch.randn(B, H, W, CI, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(CO, TAPS * CI, device="cuda") / (TAPS * CI) ** 0.5).bfloat16()
    sc, bi = torch.rand(CO, device="cuda") + 0.5, torch.randn(CO, device="cuda") * 0.1
    gap = torch.zeros(B, CO, device="cuda") if len(sys.argv) > 8 and sys.argv[8] == "gap" else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(6):
        if i == 3:
            e0.record()
        nat.conv_gemm(x, w, taps=TAPS, scale=sc, bias=bi, act=1, gap=gap)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"conv {B}x{H}x{W} {CI}->{CO} taps={TAPS} gap={gap is not None}: {ms * 1e3:.1f} us  {2.0 * B * H * W * CI * CO * TAPS / ms / 1e9:.0f} TFLOP/s")
    sys.exit(0)
M, K, N = (int(a) for a in sys.argv[1:4])
kind = sys.argv[4] if len(sys.argv) > 4 else "gelu"
dev = "cuda"
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
bias = torch.randn(N, device=dev) * 0.1
res = torch.randn(M, N, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(6):
    if i == 3:
        e0.record()
    if kind == "gelu":
        nat.linear(x, w, bias=bias, act=1)
    else:
        nat.linear_f32(x, w, bias=bias, res=res, res_mode=2, out_dtype=torch.float32)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{kind} M={M} K={K} N={N}: {ms * 1e3:.1f} us  {2.0 * M * K * N / ms / 1e9:.0f} TFLOP/s")
