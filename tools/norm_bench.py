"""Time the normaliser / resize kernels alone (CUDA events, L2 flushed between iterations) and print achieved
HBM bandwidth against the algorithmic bytes of SURVEY.md section 8(d).

    python tools/norm_bench.py [--batch 1024]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200path  # noqa: F401,E402
import b200_native as nat  # noqa: E402
import dataset as ds  # noqa: E402
import preprocess_helpers as pre  # noqa: E402


def timed(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()  # 256 MB write: evicts the 126 MB L2
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
    ap.add_argument("--sweep", action="store_true",
                    help="DWI normaliser against a device-to-device copy of the SAME number of bytes, at several batch "
                         "sizes: how much of the gap to the large-copy peak is the size of the launch")
    a = ap.parse_args()
    B = a.batch
    if a.sweep:
        peak = 6531.9
        norm = ds.DWINormalize()
        for b in (256, 1024, 4096, 16384):
            x = (torch.rand(b, 16, 64, 64, device="cuda") * 1000 + 1)
            pm = torch.empty(b * 16, device="cuda")
            ms = timed(lambda: norm.batch(x, plane_mean=pm))
            nbytes = b * 507904
            src = torch.empty(nbytes // 8, dtype=torch.float32, device="cuda").normal_()
            dst = torch.empty_like(src)
            ms_c = timed(lambda: dst.copy_(src))
            print(f"B={b:6d}  dwi_normalize {ms:7.3f} ms {nbytes / ms / 1e6:6.0f} GB/s ({100 * nbytes / ms / 1e6 / peak:4.1f} %)   "
                  f"same-bytes copy {ms_c:7.3f} ms {nbytes / ms_c / 1e6:6.0f} GB/s ({100 * nbytes / ms_c / 1e6 / peak:4.1f} %)   "
                  f"normaliser / copy = {ms_c / ms:4.2f}")
        return
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    g = torch.Generator().manual_seed(0)
    dwi = (torch.rand(B, 16, 64, 64, generator=g) * 1000 + 1).cuda()
    dce = torch.rand(B, 6, 64, 64, generator=g).cuda()
    nyul = pre.NyulStandardizer()
    nyul.fit(list(dce[:32].cpu()), num_channels=6)
    norm = ds.DWINormalize()
    pm = torch.empty(B * 16, device="cuda")
    rows = []
    ms = timed(lambda: norm.batch(dwi, plane_mean=pm))
    rows.append(("dwi_normalize 16x64x64", ms, B * 507904))          # 15 planes read + 16 written
    ms = timed(lambda: norm.fused_params(dwi))
    rows.append(("dwi statistics pass (fused first layer)", ms, B * 15 * 64 * 64 * 4))   # 15 planes read
    dnorm = ds.DCENormalize(nyul)
    ms = timed(lambda: dnorm.fused_params(dce))
    rows.append(("nyul table pass (fused first layer)", ms, B * 6 * 64 * 64 * 4))
    pm6 = torch.empty(B * 6, device="cuda")
    ms = timed(lambda: nyul.transform_batch(dce, plane_mean=pm6))
    rows.append(("nyul_transform 6x64x64", ms, B * 196608))
    b4 = max(1, B // 4)
    rz = ds.Resize(224)
    ms = timed(lambda: rz.batch(dwi[:b4]))
    rows.append((f"resize 16x64^2->224^2 (B={b4})", ms, b4 * 3473408))
    d224, c224 = rz.batch(dwi[:b4]), rz.batch(dce[:b4])
    pmb = torch.empty(b4 * 16, device="cuda")
    ms = timed(lambda: norm.batch(d224, plane_mean=pmb))
    rows.append((f"dwi_normalize 16x224x224 (B={b4})", ms, b4 * (15 + 16) * 224 * 224 * 4))
    nyul224 = pre.NyulStandardizer()
    nyul224.fit(list(c224[:8].cpu()), num_channels=6)
    ms = timed(lambda: nyul224.transform_batch(c224))
    rows.append((f"nyul_transform 6x224x224 (B={b4})", ms, b4 * 2 * 6 * 224 * 224 * 4))
    peak = peaks.get("hbm_gbps") or peaks.get("hbm_copy_gbps") or 6531.9
    for name, ms, nbytes in rows:
        gbs = nbytes / ms / 1e6
        print(f"{name:42s} {ms:8.3f} ms  {gbs:8.0f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f} GB/s")


if __name__ == "__main__":
    main()
