"""Time the encoders' first layer (b200_stem_ex: modality SE + the two strided 1x1 convolutions of block 1) alone, with
the normalisation fused into its operand load as the product runs it (CUDA events, L2 flushed between iterations).

    python tools/stem_bench.py [--batch 1024]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200path  # noqa: F401,E402
import b200_native as nat  # noqa: E402


def timed(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    B, dev = a.batch, "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    for name, C in (("DWI 16x64x64", 16), ("DCE 6x64x64", 6)):
        n_skip, n_mid, cm = 128, 64, max(C // 2, 1)
        x = torch.rand(B, C, 64, 64, generator=g, device=dev)
        pm = x.mean(dim=(2, 3)).reshape(-1).contiguous()
        se = tuple(torch.randn(s, generator=g, device=dev) * 0.1 for s in ((cm, C), (cm,), (C, cm), (C,)))
        w = torch.randn(n_skip + n_mid, C, generator=g, device=dev) * 0.2
        sc, bi = torch.ones(n_skip + n_mid, device=dev), torch.zeros(n_skip + n_mid, device=dev)
        skip = torch.empty(B, 32, 32, n_skip, dtype=torch.bfloat16, device=dev)
        mid = torch.empty(B, 32, 32, n_mid, dtype=torch.bfloat16, device=dev)
        attn = torch.empty(B, C, device=dev)
        ms = timed(lambda: nat.stem(x, 2, pm, se, w, sc, bi, n_skip, n_mid, skip, mid, attn), a.iters)
        nbytes = B * (C * 32 * 64 * 4 + 1024 * (n_skip + n_mid) * 2)  # every other input row, both output maps
        print(f"stem {name:14s} B={B}  {ms:.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
