"""Kernel-level numerics on the GPU: every C-ABI op against a plain PyTorch fp32 restatement
of the same op on the same (bf16-rounded) operands.  Tolerances are stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

import b200_native as nat

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-12)


def _conv_ref(x, w, taps, scale, bias, res, res_mode, act):
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    xf = x.float().permute(0, 3, 1, 2)
    k = 3 if taps == 9 else 1
    wf = w.float().view(Cout, taps, Cin).permute(0, 2, 1).reshape(Cout, Cin, k, k)
    y = F.conv2d(xf, wf, padding=k // 2)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    r = res.float().permute(0, 3, 1, 2) if res is not None else None
    if res_mode == 1:
        y = y + r
    if act:
        y = F.gelu(y)
    if res_mode == 2:
        y = y + r
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps,res_mode,act,gap,up2", [
    (2, 32, 32, 64, 64, 1, 0, 0, False, False),
    (2, 32, 32, 64, 64, 9, 0, 1, False, False),
    (3, 32, 32, 64, 128, 1, 1, 1, True, False),
    (2, 32, 32, 128, 128, 9, 0, 1, False, False),
    (2, 32, 32, 128, 256, 1, 2, 1, False, False),
    (5, 32, 32, 256, 256, 9, 0, 1, True, False),
    (2, 32, 32, 256, 512, 1, 1, 1, True, False),
    (2, 32, 32, 128, 64, 1, 0, 1, False, True),
    (1, 64, 64, 64, 64, 9, 0, 1, False, False),
    (40, 32, 32, 256, 256, 9, 0, 1, True, False),   # > 148 tiles: persistent loop + both TMEM stages
])
def test_conv_gemm(B, H, W, Cin, Cout, taps, res_mode, act, gap, up2):
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + Cin + Cout + taps)
    x = (torch.randn(B, H, W, Cin, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(Cout, taps * Cin, generator=g) / math.sqrt(taps * Cin)).to(DEV).bfloat16()
    scale = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
    bias = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    res = (torch.randn(B, H, W, Cout, generator=g) * 0.5).to(DEV).bfloat16() if res_mode else None
    gap_buf = torch.zeros(B, Cout, device=DEV) if gap else None
    y = nat.conv_gemm(x, w, taps=taps, scale=scale, bias=bias, res=res, res_mode=res_mode, act=act, gap=gap_buf,
                      up2=up2)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, taps, scale, bias, res, res_mode, act)
    if up2:
        ref = ref.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    # bf16 output rounding: 2^-9 relative per element; accumulation is fp32
    assert _rel(y, ref) < 1e-2
    assert (y.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    if gap:
        gref = _conv_ref(x, w, taps, scale, bias, res, res_mode, act).sum(dim=(1, 2))
        assert _rel(gap_buf, gref) < 1e-3


@pytest.mark.parametrize("M,K,N", [(197, 768, 768), (1000, 128, 64), (128, 64, 256), (77, 3072, 768)])
def test_linear_ragged_rows(M, K, N):
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    x = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    bias = (torch.randn(N, generator=g) * 0.1).to(DEV)
    y = nat.linear(x, w, bias=bias, act=1)
    torch.cuda.synchronize()
    ref = F.gelu(x.float() @ w.float().t() + bias)
    assert _rel(y, ref) < 1e-2


@pytest.mark.parametrize("B,Cin,n1,n2", [(3, 128, 256, 128), (2, 256, 512, 256), (2, 64, 64, 64)])
def test_conv_gemm_two_segments(B, Cin, n1, n2):
    """Skip conv (no activation) and first bottleneck conv (GELU) on the same input in one launch."""
    g = torch.Generator(device="cpu").manual_seed(Cin + n1)
    x = (torch.randn(B, 32, 32, Cin, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(n1 + n2, Cin, generator=g) / math.sqrt(Cin)).to(DEV).bfloat16()
    scale = (torch.rand(n1 + n2, generator=g) + 0.5).to(DEV)
    bias = (torch.randn(n1 + n2, generator=g) * 0.1).to(DEV)
    y1, y2 = nat.conv_gemm(x, w, taps=1, scale=scale, bias=bias, act=0, n_split=n1, act2=1)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, 1, scale, bias, None, 0, 0)
    assert y1.shape[-1] == n1 and y2.shape[-1] == n2
    assert _rel(y1, ref[..., :n1]) < 1e-2
    assert _rel(y2, F.gelu(ref[..., n1:])) < 1e-2


@pytest.mark.parametrize("B,C", [(2, 128), (3, 256), (1, 64)])
def test_recon_head_fused_tap_dots(B, C):
    """3x3 conv + affine + GELU with the following 3x3, C->1 conv folded into the epilogue + tapsum."""
    g = torch.Generator(device="cpu").manual_seed(C)
    x = (torch.randn(B, 32, 32, C, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(C, 9 * C, generator=g) / math.sqrt(9 * C)).to(DEV).bfloat16()
    scale = (torch.rand(C, generator=g) + 0.5).to(DEV)
    bias = (torch.randn(C, generator=g) * 0.1).to(DEV)
    w3 = (torch.randn(9, C, generator=g) / math.sqrt(9 * C)).to(DEV)
    b3 = torch.tensor([0.3], device=DEV)
    d = torch.empty(B, 32, 32, 9, device=DEV)
    nat.conv_gemm(x, w, taps=9, scale=scale, bias=bias, act=1, store=False, dot_w=w3, dot_out=d)
    out = nat.tapsum(d, b3, torch.empty(B, 32, 32, device=DEV))
    torch.cuda.synchronize()
    t = _conv_ref(x, w, 9, scale, bias, None, 0, 1)  # fp32, not rounded to bf16 (the fused path keeps fp32 too)
    ref = F.conv2d(t.permute(0, 3, 1, 2), w3.view(9, C).t().reshape(1, C, 3, 3), b3, padding=1)[:, 0]
    assert _rel(out, ref) < 5e-3
    # and the stand-alone N=1 kernel on a bf16 map agrees with torch as well
    tb = t.bfloat16()
    out2 = nat.conv3x3_c1(tb, w3, b3, torch.empty(B, 32, 32, device=DEV))
    ref2 = F.conv2d(tb.float().permute(0, 3, 1, 2), w3.view(9, C).t().reshape(1, C, 3, 3), b3, padding=1)[:, 0]
    assert _rel(out2, ref2) < 1e-3


def test_fast_gelu_deviation_is_below_bf16_resolution():
    """The epilogue GELU uses an 8-term odd polynomial for erf: compare with exact GELU through an
    identity 1x1 'convolution' (weights = I)."""
    C = 64
    xs = torch.linspace(-9, 9, 2 * 32 * 32 * C).view(2, 32, 32, C).to(DEV).bfloat16()
    w = torch.eye(C, device=DEV).bfloat16()
    y = nat.conv_gemm(xs, w, taps=1, act=1)
    torch.cuda.synchronize()
    ref = F.gelu(xs.float())
    err = (y.float() - ref).abs()
    assert (err <= 4e-3 * ref.abs() + 2.5e-4).all(), err.max()


def test_fused_single_dot_is_a_following_1x1_conv_to_one_channel():
    """Mask head: `pre` 1x1 conv + bias with `out` (C -> 1) folded into the epilogue (ndot = 1)."""
    g = torch.Generator(device="cpu").manual_seed(3)
    B, Cin, C = 3, 256, 64
    x = (torch.randn(B, 32, 32, Cin, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(C, Cin, generator=g) / math.sqrt(Cin)).to(DEV).bfloat16()
    bias = (torch.randn(C, generator=g) * 0.1).to(DEV)
    wo = (torch.randn(1, C, generator=g) / math.sqrt(C)).to(DEV)
    out = torch.empty(B, 32, 32, 1, device=DEV)
    nat.conv_gemm(x, w, taps=1, bias=bias, store=False, dot_w=wo, dot_out=out, dot_bias=0.25)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t() + bias) @ wo.t() + 0.25
    assert _rel(out, ref) < 1e-4
