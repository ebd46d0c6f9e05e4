"""GPU: the training-mode kernels (csrc/conv_wgrad.cu, csrc/train_elem.cu) against torch fp32 restatements of the same
ops, and the whole single-modality training step (BASELINE config C1) against the CPU oracle and the fixture the
unmodified reference produced (tests/golden/train_c1_dwi.npz)."""
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import b200path  # noqa: F401
import golden_util as gu
from oracle import params as op
from oracle import train_oracle as to

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() / max(b.double().abs().max().item(), 1e-12)


def _nhwc_bf16(t):  # NCHW fp32 -> NHWC bf16 device
    return t.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)


@pytest.mark.parametrize("cin,cout,taps,hw", [(64, 64, 9, 32), (128, 128, 9, 32), (256, 256, 9, 32), (128, 256, 1, 32),
                                              (256, 512, 1, 32), (512, 128, 1, 32), (64, 64, 1, 64), (128, 64, 1, 16)])
def test_conv_wgrad_and_dgrad_vs_torch(cin, cout, taps, hw):
    import b200_native as nat

    g = torch.Generator().manual_seed(cin * 7 + cout + taps)
    B = 3
    x = torch.randn(B, cin, hw, hw, generator=g).bfloat16().float()
    dy = torch.randn(B, cout, hw, hw, generator=g).bfloat16().float()
    k = 3 if taps == 9 else 1
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5)
    ref_dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=k // 2)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.bfloat16().float(), dy, padding=k // 2)
    xd, dyd = _nhwc_bf16(x), _nhwc_bf16(dy)
    dw = torch.zeros_like(w, device=DEV)
    nat._call("b200_conv_wgrad", None, dyd.data_ptr(), cout, xd.data_ptr(), cin, dw.data_ptr(), B, hw, hw, cin, cout, taps,
              nat._stream())
    assert _rel(dw, ref_dw) <= 2e-3, f"wgrad {_rel(dw, ref_dw):.2e}"
    # accumulation: a second launch doubles the result
    nat._call("b200_conv_wgrad", None, dyd.data_ptr(), cout, xd.data_ptr(), cin, dw.data_ptr(), B, hw, hw, cin, cout, taps,
              nat._stream())
    assert _rel(dw, 2 * ref_dw) <= 2e-3
    wd_dev = w.to(DEV)
    wf = torch.empty((cout, taps * cin), dtype=torch.bfloat16, device=DEV)
    wd = torch.empty((cin, taps * cout), dtype=torch.bfloat16, device=DEV)
    nat._call("b200_pack_conv_weights", None, wd_dev.data_ptr(), cout, cin, taps, wf.data_ptr(), wd.data_ptr(), nat._stream())
    assert torch.equal(wf.cpu(), w.permute(0, 2, 3, 1).reshape(cout, -1).bfloat16())
    dx = nat.conv_gemm(dyd, wd, taps=taps)
    assert _rel(dx.float().permute(0, 3, 1, 2), ref_dx) <= 1e-2, f"dgrad {_rel(dx.float().permute(0, 3, 1, 2), ref_dx):.2e}"


@pytest.mark.parametrize("C,act,with_res,drop", [(64, 1, False, 0.0), (256, 1, True, 0.0), (128, 0, False, 0.0),
                                                 (512, 1, True, 0.2)])
def test_bn_act_forward_backward_vs_torch(C, act, with_res, drop):
    import torch.nn as nn

    import train_graph as tg

    g = torch.Generator().manual_seed(C + act)
    B, H, W = 4, 16, 16
    z = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).bfloat16().float()
    res = torch.randn(B, C, H, W, generator=g).bfloat16().float() if with_res else None
    da = torch.randn(B, C, H, W, generator=g).bfloat16().float()
    bn = nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g))
        bn.bias.copy_(0.1 * torch.randn(C, generator=g))
    bn_dev = nn.BatchNorm2d(C).to(DEV)
    bn_dev.load_state_dict(bn.state_dict())
    ops = tg.TrainOps(torch.device(DEV))
    zd = _nhwc_bf16(z)
    rd = _nhwc_bf16(res) if with_res else None
    a = ops.bn_act(zd, bn=bn_dev, act=act, res=rd, drop_p=drop)
    ops.tape.add_grad(a, _nhwc_bf16(da))
    a_nchw = a.float().permute(0, 3, 1, 2).cpu()
    if drop > 0:  # the dropout mask is the kernel's own: recover it from the output, then check everything else given it
        zr = z.clone().requires_grad_(True)
        y = bn(zr) + (res if with_res else 0)
        y = F.gelu(y) if act == 1 else y
        keep = (a_nchw != 0) | (y.detach().abs() < 1e-3)
        rate = 1 - keep.float().mean().item()
        assert abs(rate - drop) < 0.02, rate
        ref = y * keep / (1 - drop)
    else:
        zr = z.clone().requires_grad_(True)
        rr = res.clone().requires_grad_(True) if with_res else None
        y = bn(zr) + (rr if with_res else 0)
        ref = F.gelu(y) if act == 1 else y
    assert _rel(a_nchw, ref.detach()) <= 8e-3
    ref.backward(da)
    ops.tape.backward()
    # (the tape consumed the gradients: fetch what the kernels accumulated)
    tol = 3e-2 if drop > 0 else 1e-2  # (with dropout the mask is recovered from the bf16 output: a few ambiguous zeros)
    assert _rel(bn_dev.weight.grad, bn.weight.grad) <= tol
    assert _rel(bn_dev.bias.grad, bn.bias.grad) <= tol
    assert _rel(bn_dev.running_mean, bn.running_mean) <= 1e-4 and _rel(bn_dev.running_var, bn.running_var) <= 1e-3


def test_bn_act_backward_map_gradients_vs_torch():
    """dz and dres of the fused BN + residual + GELU backward, read back through a probe step on the tape."""
    import torch.nn as nn

    import train_graph as tg

    g = torch.Generator().manual_seed(5)
    B, C, H, W = 4, 128, 16, 16
    z = (torch.randn(B, C, H, W, generator=g) * 2).bfloat16().float()
    res = torch.randn(B, C, H, W, generator=g).bfloat16().float()
    da = torch.randn(B, C, H, W, generator=g).bfloat16().float()
    bn = nn.BatchNorm2d(C)
    bn_dev = nn.BatchNorm2d(C).to(DEV)
    ops = tg.TrainOps(torch.device(DEV))
    zd, rd = _nhwc_bf16(z), _nhwc_bf16(res)
    got = {}
    ops.tape.record(lambda: got.update(dz=ops.tape.grad_of(zd), dres=ops.tape.grad_of(rd)))
    a = ops.bn_act(zd, bn=bn_dev, act=1, res=rd)
    ops.tape.add_grad(a, _nhwc_bf16(da))
    ops.tape.backward()
    zr, rr = z.clone().requires_grad_(True), res.clone().requires_grad_(True)
    F.gelu(bn(zr) + rr).backward(da)
    assert _rel(got["dz"].float().permute(0, 3, 1, 2), zr.grad) <= 1.5e-2
    assert _rel(got["dres"].float().permute(0, 3, 1, 2), rr.grad) <= 1e-2


def _c1_setup(n):
    import model_module as mm
    import parameters_default as pd

    gold = gu.load("train_c1_dwi.npz")
    hp = json.loads(str(gold["hp"]))
    p = pd.default_parameters()
    p["dwi_model_parameters"]["dropout"] = 0.0
    model = mm.ModelMaskHeadBackbone("dwi", p)
    sd = op.seeded_state_dict(op.shapes_of(model.state_dict()), seed=hp["weight_seed"])
    model.load_state_dict(sd)
    model.to(DEV).train()
    for q in model.parameters():
        q.requires_grad_(True)
    dwi_raw, _, masks, labels = op.synthetic_raw(n, seed=1234, kind="S")
    x = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    return gold, hp, p, model, sd, x, masks, labels


def test_c1_single_modality_train_step_vs_oracle_and_reference_fixture():
    """BASELINE config C1 on the GPU: forward + loss + backward of the DWI CNN encoder in train mode (batch-statistic
    BatchNorm; dropouts at p = 0 as in the fixture).  Loss terms and all parameter gradients against the CPU oracle and
    the fixture produced by the reference's own modules and loss functions."""
    import train_graph as tg

    gold, hp, p, model, sd, x, masks, labels = _c1_setup(8)
    lam = {k: hp[k] for k in ("lambda_mask", "lambda_recon", "lambda_mimic", "lambda_feat_norm")}
    cw = torch.tensor(hp["class_weights"])
    ops = tg.TrainOps(torch.device(DEV))
    out = tg.encoder_forward_train(ops, model, x.to(DEV))
    total, parts = tg.single_model_loss(ops, out, masks.to(DEV), labels.to(DEV), smoothing=hp["smoothing"],
                                        gamma=hp["gamma"], class_weights=cw, **lam)
    ops.tape.backward()
    torch.cuda.synchronize()
    o_total, o_parts, o_grads = to.single_model_objective_and_grads(sd, p, "dwi", x, masks, labels, hp["smoothing"],
                                                                    hp["gamma"], cw, **lam)
    rep = {k: (parts[k].item(), o_parts[k]) for k in o_parts}
    print("loss parts (gpu, oracle):", rep, "total", total.item(), float(o_total))
    gp = gold["parts"]  # total, cls, feat_norm, mask, recon_w, mimic_w (fixture; the last two already lambda-weighted)
    assert abs(total.item() - gp[0]) <= 1e-2 * abs(gp[0]), (total.item(), gp[0])
    assert abs(total.item() - float(o_total)) <= 1e-2 * abs(float(o_total))
    for k in o_parts:
        assert abs(parts[k].item() - o_parts[k]) <= 2e-2 * max(abs(o_parts[k]), 1e-3), (k, parts[k].item(), o_parts[k])
    named = dict(model.named_parameters())
    errs, missing = {}, []
    for k, g in o_grads.items():
        got = named[k].grad
        if got is None:
            missing.append(k)
            continue
        errs[k] = _rel(got, g)
    assert not missing, missing
    extra = [k for k, q in named.items() if q.grad is not None and k not in o_grads and q.grad.abs().max().item() > 0]
    assert not extra, extra
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print("worst gradient errors:", [(k, f"{v:.2e}") for k, v in worst[:10]])
    assert len(errs) == len(hp["with_grad"]) == 89
    # bf16 activations and bf16 activation gradients against an fp32 oracle.  The six tiny parameters of the mask-guided
    # attention (a scalar gamma, 16-vectors) are sums over every pixel and channel of products of bf16 gradient maps that
    # cancel almost completely: they carry the rounding noise of those maps (measured 1 - 32 % run to run); everything else - all
    # convolution, BatchNorm, squeeze-excite, head and projector gradients - is held to 4e-2, the median to 1.5e-2.
    noisy = {k: v for k, v in errs.items() if k.startswith("mask_spatial_attention.")}
    rest = {k: v for k, v in errs.items() if k not in noisy}
    assert max(rest.values()) <= 4e-2, sorted(rest.items(), key=lambda kv: -kv[1])[:5]
    assert max(noisy.values()) <= 5e-1 and np.median(list(noisy.values())) <= 1.5e-1, noisy  # noise-level sanity bound
    assert np.median([v for _, v in worst]) <= 1.5e-2
    # ... and directly against the reference fixture (full tensors where the fixture stores them, i.e. <= 8 192 elements)
    n_direct = 0
    for k in hp["with_grad"]:
        if k in noisy or f"grad/{k}/full" not in gold.files:
            continue
        assert _rel(named[k].grad, torch.from_numpy(gold[f"grad/{k}/full"])) <= 4e-2, k
        n_direct += 1
    assert n_direct >= 30
