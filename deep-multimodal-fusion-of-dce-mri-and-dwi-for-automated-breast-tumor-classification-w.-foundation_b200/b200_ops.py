"""torch.library registration of the path's batched entry points (`torch.ops.b200.*`).

The Python modules of this package call the C ABI through `b200_native` directly; this module registers the same
entry points as PyTorch custom operators with fake (meta) implementations, so that graph capture - `torch.compile`,
which the reference wraps its models in when `parameters['compile']` is set (code/run_training.py:90-91, :242-244),
`make_fx`, export - sees opaque, shape-correct nodes instead of ctypes calls.  The real implementations are the
hand-written sm_100a kernels (no CPU implementation is registered: a CPU tensor raises `B200NativeError`).

    import b200_ops                      # registers the operators
    out, plane_mean = torch.ops.b200.dwi_normalize(x, True, -3.0, 3.0)

Operators (all functional, outputs freshly allocated):
  dwi_normalize(x[B,C,H,W] f32, adc, z_lo, z_hi) -> (out[B,C,H,W] f32, plane_mean[B*C] f32)     code/dataset.py:14-41
  nyul_transform(x[B,C,H,W] f32, avg[C,L] f64, scale[L] f64, prev[..] i32, gamma[..] f64) -> (out, plane_mean)
                                                                                     code/preprocess_helpers.py:85-120
  resize_bilinear(x[B,C,h,w] f32, H, W) -> [B,C,H,W] f32                             code/prepare_single_model.py:112-120
  augment(x[B,C,H,W] f32, theta[B,6] f32, flips[B] i32, fill) -> [B,C,H,W] f32       code/prepare_single_model.py:107-113
  conv_gemm(x[B,H,W,Cin] bf16, w[Cout,taps*Cin] bf16, scale[Cout]?, bias[Cout]?, res?, res_mode, act, taps)
      -> [B,H,W,Cout] bf16                                                           code/model_module.py:259-269 etc.
  fusion_tokens(p[B,H,W,C] bf16, hp, wp) -> [B,hp*wp,C] f32                          code/model_module.py:903-917
  attention(qkv[B*N,3*heads*d] bf16, B, N, heads, d) -> [B*N,heads*d] bf16           code/transformer_model.py:101-112
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.library import custom_op

import b200_native as nat

__all__ = ["dwi_normalize", "nyul_transform", "resize_bilinear", "augment", "conv_gemm", "fusion_tokens", "attention"]


def _need_cuda(t, name):
    if not t.is_cuda:
        raise nat.B200NativeError(f"b200::{name} needs CUDA tensors (there is no CPU implementation)")


@custom_op("b200::dwi_normalize", mutates_args=())
def dwi_normalize(x: torch.Tensor, adc: bool, z_lo: float, z_hi: float) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(x, "dwi_normalize")
    x = x.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    pm = torch.empty(B * C, dtype=torch.float32, device=x.device)
    nat.dwi_normalize(x, out, C, H * W, adc, z_lo, z_hi, pm)
    return out, pm


@dwi_normalize.register_fake
def _(x, adc, z_lo, z_hi):
    B, C, H, W = x.shape
    return x.new_empty((B, C, H, W), dtype=torch.float32), x.new_empty((B * C,), dtype=torch.float32)


@custom_op("b200::nyul_transform", mutates_args=())
def nyul_transform(x: torch.Tensor, avg_landmarks: torch.Tensor, standard_scale: torch.Tensor,
                   prev_index: torch.Tensor, gamma: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(x, "nyul_transform")
    x = x.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    pm = torch.empty(B * C, dtype=torch.float32, device=x.device)
    nat.nyul_transform(x, out, C, H * W, avg_landmarks, standard_scale, prev_index, gamma, pm)
    return out, pm


@nyul_transform.register_fake
def _(x, avg_landmarks, standard_scale, prev_index, gamma):
    B, C, H, W = x.shape
    return x.new_empty((B, C, H, W), dtype=torch.float32), x.new_empty((B * C,), dtype=torch.float32)


@custom_op("b200::resize_bilinear", mutates_args=())
def resize_bilinear(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    _need_cuda(x, "resize_bilinear")
    x = x.contiguous().float()
    B, C, h, w = x.shape
    out = torch.empty((B, C, height, width), dtype=torch.float32, device=x.device)
    # a shrinking side takes the antialiased triangle filter, as torchvision's Resize does for tensors
    kernel = nat.resize_aa_c1 if (height < h or width < w) else nat.resize_bilinear_c1
    kernel(x.view(B * C, h, w), out.view(B * C, height, width))
    return out


@resize_bilinear.register_fake
def _(x, height, width):
    return x.new_empty((x.shape[0], x.shape[1], height, width), dtype=torch.float32)


@custom_op("b200::augment", mutates_args=())
def augment(x: torch.Tensor, theta: torch.Tensor, flips: torch.Tensor, fill: float) -> torch.Tensor:
    _need_cuda(x, "augment")
    x = x.contiguous().float()
    out = torch.empty_like(x)
    nat.augment(x, theta.contiguous().float(), flips.contiguous().to(torch.int32), out, fill)
    return out


@augment.register_fake
def _(x, theta, flips, fill):
    return x.new_empty(x.shape, dtype=torch.float32)


@custom_op("b200::conv_gemm", mutates_args=())
def conv_gemm(x: torch.Tensor, w: torch.Tensor, scale: Optional[torch.Tensor], bias: Optional[torch.Tensor],
              res: Optional[torch.Tensor], res_mode: int, act: int, taps: int) -> torch.Tensor:
    _need_cuda(x, "conv_gemm")
    return nat.conv_gemm(x, w, taps=taps, scale=scale, bias=bias, res=res, res_mode=res_mode, act=act)


@conv_gemm.register_fake
def _(x, w, scale, bias, res, res_mode, act, taps):
    B, H, W, _ = x.shape
    return x.new_empty((B, H, W, w.shape[0]), dtype=torch.bfloat16)


@custom_op("b200::fusion_tokens", mutates_args=())
def fusion_tokens(p: torch.Tensor, hp: int, wp: int) -> torch.Tensor:
    _need_cuda(p, "fusion_tokens")
    B, H, W, C = p.shape
    tokens = torch.empty((B, hp * wp, C), dtype=torch.float32, device=p.device)
    nat.fusion_tokens(p.contiguous(), hp, wp, tokens)
    return tokens


@fusion_tokens.register_fake
def _(p, hp, wp):
    return p.new_empty((p.shape[0], hp * wp, p.shape[3]), dtype=torch.float32)


@custom_op("b200::attention", mutates_args=())
def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int) -> torch.Tensor:
    """softmax(q k^T / sqrt d) v per (case, head) from the packed rows `Linear(E, 3E)` writes (b200_attention)."""
    _need_cuda(qkv, "attention")
    qkv = qkv.contiguous()
    out = torch.empty((batch * tokens, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    nat.attention(qkv, out, batch, tokens, heads, head_dim)
    return out


@attention.register_fake
def _(qkv, batch, tokens, heads, head_dim):
    return qkv.new_empty((batch * tokens, heads * head_dim), dtype=torch.bfloat16)
