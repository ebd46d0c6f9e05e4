"""Times the training kernels one by one on the shapes of the C5 step (CUDA events, L2 flushed between iterations) and
prints achieved HBM bandwidth (algorithmic bytes) or tensor throughput.

    python tools/train_kbench.py [--batch 512]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200path  # noqa: F401,E402
import b200_native as nat  # noqa: E402

P = nat._ptr


def timed(fn, iters=5):
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


def call(name, *a):
    nat._call(name, None, *a, nat._stream())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    B = ap.parse_args().batch
    dev = "cuda"
    H = W = 32
    R = B * H * W
    rows = []

    def bf(*shape):
        return torch.randn(shape, device=dev).bfloat16()

    def f32(*shape):
        return torch.randn(shape, device=dev)

    scratch = torch.zeros(4096, dtype=torch.float64, device=dev)
    for C in (64, 128, 256, 512):
        z, res, dA, out, dz, dres = bf(R, C), bf(R, C), bf(R, C), bf(R, C), bf(R, C), bf(R, C)
        mean, inv, g, b_ = f32(C), f32(C).abs() + 0.5, f32(C), f32(C)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        ms = timed(lambda: call("b200_bn_stats", P(z), R, C, C, P(scratch), P(scratch[C:])))
        rows.append((f"bn_stats C={C}", ms, R * C * 2))
        ms = timed(lambda: call("b200_bn_act_fwd", P(z), C, P(res), C, P(mean), P(inv), P(g), P(b_), 1, 0.2, 1234, R, C, P(out), C))
        rows.append((f"bn_act_fwd C={C} (+res, GELU, dropout)", ms, R * C * 2 * 3))
        ms = timed(lambda: call("b200_bn_act_fwd", P(z), C, None, 0, P(mean), P(inv), P(g), P(b_), 1, 0.0, 0, R, C, P(out), C))
        rows.append((f"bn_act_fwd C={C} (GELU)", ms, R * C * 2 * 2))
        ms = timed(lambda: call("b200_bn_act_bwd", P(z), C, P(res), C, P(mean), P(inv), P(g), P(b_), 1, 0.2, 1234, R, C, P(dA), C,
                                1, P(scratch), P(dz), C, P(dres), C, P(dg), P(db)))
        rows.append((f"bn_act_bwd C={C} (+res, GELU, dropout; 2 passes)", ms, R * C * 2 * (3 + 2 + 3)))
        ms = timed(lambda: call("b200_bn_act_bwd", P(z), C, None, 0, P(mean), P(inv), P(g), P(b_), 1, 0.0, 0, R, C, P(dA), C,
                                1, P(scratch), P(dz), C, None, 0, P(dg), P(db)))
        rows.append((f"bn_act_bwd C={C} (GELU; 2 passes)", ms, R * C * 2 * (2 + 1 + 3)))
        a, o = bf(R, C), torch.empty(B, C, device=dev)
        ms = timed(lambda: call("b200_map_dot", P(a), C, P(z), C, B, H * W, C, P(o)))
        rows.append((f"map_dot C={C}", ms, R * C * 2 * 2))
        gate = f32(B, C)
        ms = timed(lambda: call("b200_map_scale_add", P(a), C, P(gate), None, B, H * W, C, P(out), C, 0))
        rows.append((f"map_scale_add C={C}", ms, R * C * 2 * 2))
    for C, taps in ((128, 9), (256, 9), (64, 1)):
        x, w, bias = bf(B, H, W, C), f32(C * taps), f32(1)
        o, do = torch.empty(B, H, W, device=dev), f32(B, H, W)
        dx, dw, dbias = bf(B, H, W, C), torch.zeros(C * taps, device=dev), torch.zeros(1, device=dev)
        ms = timed(lambda: call("b200_convc1_fwd", P(x), C, B, H, W, C, taps, P(w), P(bias), P(o)))
        rows.append((f"convc1_fwd C={C} taps={taps}", ms, R * C * 2))
        ms = timed(lambda: call("b200_convc1_bwd", P(x), C, P(do), B, H, W, C, taps, P(w), P(dx), C, 0, None, None))
        rows.append((f"convc1_bwd dx C={C} taps={taps}", ms, R * C * 2))
        ms = timed(lambda: call("b200_convc1_bwd", P(x), C, P(do), B, H, W, C, taps, P(w), None, 0, 0, P(dw), P(dbias)))
        rows.append((f"convc1_bwd dw C={C} taps={taps}", ms, R * C * 2))
    print(f"--- HBM-bound kernels, B = {B} (peak 6532 GB/s) ---")
    for name, ms, nbytes in rows:
        print(f"{name:52s} {ms:8.3f} ms {nbytes / ms / 1e6:8.0f} GB/s {100 * nbytes / ms / 1e6 / 6531.9:5.1f} %")
    print(f"--- tensor-core kernels, B = {B} (sustained peak 1421 TF/s) ---")
    for cin, cout, taps in ((64, 64, 9), (128, 128, 9), (256, 256, 9), (128, 256, 1), (256, 512, 1), (256, 256, 1), (512, 128, 1),
                            (128, 64, 1), (256, 64, 1), (64, 64, 1)):
        x, dy = bf(B, H, W, cin), bf(B, H, W, cout)
        dw = torch.zeros(cout, cin, taps, device=dev)
        wf, wd = bf(cout, taps * cin), bf(cin, taps * cout)
        fl = 2.0 * R * cin * cout * taps
        ms = timed(lambda: call("b200_conv_wgrad", P(dy), cout, P(x), cin, P(dw), B, H, W, cin, cout, taps))
        ms_f = timed(lambda: nat.conv_gemm(x, wf, taps=taps))
        ms_d = timed(lambda: nat.conv_gemm(dy, wd, taps=taps))
        print(f"conv {cin:3d}->{cout:3d} taps {taps}: wgrad {ms:7.3f} ms {fl / ms / 1e9:7.0f} TF/s | fwd {ms_f:7.3f} ms "
              f"{fl / ms_f / 1e9:7.0f} TF/s | dgrad {ms_d:7.3f} ms {fl / ms_d / 1e9:7.0f} TF/s")
    x = torch.rand(B, 16, 64, 64, device=dev)
    gate, dz, wcat = torch.rand(B, 16, device=dev), bf(B, 32, 32, 192), f32(192, 16)
    dwc, dgt = torch.zeros(192, 16, device=dev), torch.zeros(B, 16, device=dev)
    ms = timed(lambda: call("b200_stem_bwd", P(x), B, 16, 64, 64, 2, P(gate), P(dz), 192, P(wcat), P(dwc), P(dgt)))
    print(f"stem_bwd: {ms:.3f} ms")


if __name__ == "__main__":
    main()
