"""Fusion-head fine-tuning step on the GPU: every training kernel against a plain PyTorch fp32 restatement, the
whole step against the CPU oracle (oracle/train_oracle.py) and the reference fixture (tests/golden/train_head.npz).
fp32 kernels: tolerances are fp32 summation-order tolerances, stated per test."""
import json

import pytest
import torch
import torch.nn.functional as F

import b200_native as nat
import golden_util as gu
from oracle import params as op
from oracle import train_oracle as to
from test_train_cpu import check_updated_parameter

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@pytest.mark.parametrize("M,N,K,ta,tb", [(64, 64, 64, 0, 0), (100, 37, 53, 0, 1), (37, 130, 300, 1, 0),
                                         (16384, 128, 512, 0, 1), (128, 512, 4096, 1, 0), (2500, 70, 129, 1, 1),
                                         (4, 258, 1000, 1, 0), (13000, 200, 77, 0, 0), (12900, 130, 100, 1, 1),
                                         (40000, 256, 128, 0, 1), (39000, 257, 64, 1, 0)])
def test_sgemm_matches_matmul(M, N, K, ta, tb):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    a = torch.randn((K, M) if ta else (M, K), generator=g, device=DEV)
    b = torch.randn((N, K) if tb else (K, N), generator=g, device=DEV)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    out = torch.empty(M, N, device=DEV)
    nat.sgemm(a, b, out, trans_a=bool(ta), trans_b=bool(tb))
    assert _rel(out, ref) < 5e-6   # fp32 accumulation over up to 4 096 terms against an fp64 product
    # split-K accumulates into the existing contents
    base = torch.randn(M, N, generator=g, device=DEV)
    out2 = base.clone()
    nat.sgemm(a, b, out2, trans_a=bool(ta), trans_b=bool(tb), beta=1, split_k=max(1, min(7, K // 16)))
    assert _rel(out2, ref + base.double()) < 5e-6


def test_sgemm_epilogue_and_strided_views():
    g = torch.Generator(device=DEV).manual_seed(3)
    R, T, C = 96, 16, 128
    x = torch.randn(R, C, generator=g, device=DEV)
    w3 = torch.randn(3 * C, C, generator=g, device=DEV) * 0.1
    b3 = torch.randn(3 * C, generator=g, device=DEV)
    kv = torch.empty(R, 2 * C, device=DEV)
    nat.sgemm(x, w3[C:], kv, trans_b=True, bias=b3[C:])
    assert _rel(kv, F.linear(x, w3[C:], b3[C:])) < 2e-6
    res = torch.randn(R // T, C, generator=g, device=DEV)
    out, pre = torch.empty(R, C, device=DEV), torch.empty(R, C, device=DEV)
    nat.sgemm(kv[:, C:], w3[:C], out, bias=b3[:C], res=res, res_div=T, pre=pre, act=1)   # strided A view
    want = kv[:, C:] @ w3[:C] + b3[:C] + res.repeat_interleave(T, dim=0)
    assert _rel(pre, want) < 2e-6 and _rel(out, F.gelu(want)) < 2e-6
    acc = out.clone()
    nat.sgemm(x, w3[:C], acc, trans_b=True, beta=1)
    assert _rel(acc, out + x @ w3[:C].t()) < 2e-6
    with pytest.raises(nat.B200NativeError):
        nat.sgemm(x, w3[:C], torch.empty(R, C + 1, device=DEV), trans_b=True)
    with pytest.raises(nat.B200NativeError):   # epilogue terms are not defined under split-K
        nat.sgemm(x, w3[:C], out, trans_b=True, bias=b3[:C], beta=1, split_k=2)


def test_colsum_accumulates():
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(5000, 130, generator=g, device=DEV)
    out = torch.ones(130, device=DEV)
    nat.colsum(x, out)
    assert _rel(out, x.double().sum(0) + 1) < 1e-5
    wide = torch.randn(300, 256, generator=g, device=DEV)
    o2 = torch.zeros(128, device=DEV)
    nat.colsum(wide[:, 128:], o2)
    assert _rel(o2, wide[:, 128:].double().sum(0)) < 1e-5


@pytest.mark.parametrize("B,NH,Tq,Tk,DH", [(5, 4, 16, 16, 32), (3, 2, 7, 12, 24)])
def test_mha_forward_and_backward_match_autograd(B, NH, Tq, Tk, DH):
    g = torch.Generator(device=DEV).manual_seed(B + Tq)
    C = NH * DH
    q = torch.randn(B * Tq, C, generator=g, device=DEV)
    kv = torch.randn(B * Tk, 2 * C, generator=g, device=DEV)
    k, v = kv[:, :C], kv[:, C:]
    probs = torch.empty(B, NH, Tq, Tk, device=DEV)
    ctx = torch.empty(B * Tq, C, device=DEV)
    nat.mha_fwd(q, k, v, B, NH, probs, ctx)
    qr, kr, vr = (t.detach().clone().requires_grad_(True) for t in (q, k.contiguous(), v.contiguous()))
    split = lambda t, T: t.view(B, T, NH, DH).permute(0, 2, 1, 3)
    s = split(qr, Tq) @ split(kr, Tk).transpose(-1, -2) / DH ** 0.5
    p_ref = s.softmax(-1)
    c_ref = (p_ref @ split(vr, Tk)).permute(0, 2, 1, 3).reshape(B * Tq, C)
    assert _rel(probs, p_ref) < 1e-5 and _rel(ctx, c_ref) < 1e-5
    dctx = torch.randn(B * Tq, C, generator=g, device=DEV)
    c_ref.backward(dctx)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    nat.mha_bwd(q, k, v, probs, dctx, B, NH, dq, dkv[:, :C], dkv[:, C:])
    assert _rel(dq, qr.grad) < 2e-5 and _rel(dkv[:, :C], kr.grad) < 2e-5 and _rel(dkv[:, C:], vr.grad) < 2e-5


def test_layernorm_and_gelu_backward_match_autograd():
    g = torch.Generator(device=DEV).manual_seed(9)
    R, C = 333, 128
    x = (torch.randn(R, C, generator=g, device=DEV) * 2 + 0.5).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(C, generator=g, device=DEV)).requires_grad_(True)
    b = (0.1 * torch.randn(C, generator=g, device=DEV)).requires_grad_(True)
    y_ref = F.layer_norm(x, (C,), w, b, 1e-5)
    y, mean, rstd = torch.empty(R, C, device=DEV), torch.empty(R, device=DEV), torch.empty(R, device=DEV)
    nat.ln_fwd(x.detach(), w.detach(), b.detach(), 1e-5, y, mean, rstd)
    assert _rel(y, y_ref) < 2e-6
    dy = torch.randn(R, C, generator=g, device=DEV)
    dres = torch.randn(R, C, generator=g, device=DEV)
    y_ref.backward(dy)
    dx, dyx = torch.empty(R, C, device=DEV), torch.empty(R, C, device=DEV)
    nat.ln_bwd(x.detach(), dy, dres, w.detach(), mean, rstd, dx, dyx)
    assert _rel(dx, x.grad + dres) < 1e-5
    gw, gb = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    nat.colsum(dyx, gw)
    nat.colsum(dy, gb)
    assert _rel(gw, w.grad) < 1e-5 and _rel(gb, b.grad) < 1e-5
    h = (torch.randn(R, C, generator=g, device=DEV) * 2).requires_grad_(True)
    F.gelu(h).backward(dy)
    out = torch.empty(R, C, device=DEV)
    nat.gelu_bwd(h.detach(), dy, out)
    assert _rel(out, h.grad) < 2e-6


def test_adamw_matches_torch_optim():
    g = torch.Generator(device=DEV).manual_seed(5)
    n = 10007
    p0 = torch.randn(n, generator=g, device=DEV)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 5):
        grad = torch.randn(n, generator=g, device=DEV)
        ref.grad = grad.clone()
        opt.step()
        nat.adamw(p, grad * 4, m, v, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.05, step=step,
                  grad_scale=0.25)
        assert _rel(p, ref.detach()) < 2e-6, step


# ------------------------------------------------------------------------------------------------------------------
# the whole step
# ------------------------------------------------------------------------------------------------------------------
def _head(seed):
    import model_module as mm
    import parameters_default as pd

    params = pd.default_parameters()
    fm = mm.FusionModel(params)
    sd = op.seeded_state_dict(op.shapes_of(fm.state_dict()), seed=seed)
    fm.load_state_dict(sd)
    return params, fm.to(DEV).eval(), sd


def _to_dev(batch):
    f3d, f3c, md, mc, labels = batch
    cl = lambda t: t.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return cl(f3d), cl(f3c), md.to(DEV), mc.to(DEV), labels.to(DEV)


def test_head_step_matches_reference_fixture_and_oracle():
    from fusion_train import FusionHeadTrainer

    gold = gu.load("train_head.npz")
    hp = json.loads(str(gold["hp"]))
    params, fm, sd = _head(hp["weight_seed"])
    batch = op.synthetic_head_batch(hp["n"], seed=hp["seed"])
    tr = FusionHeadTrainer(fm, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"], weight_decay=hp["weight_decay"],
                           smoothing=hp["smoothing"], gamma=hp["gamma"], class_weights=hp["class_weights"])
    assert sorted(tr.names) == sorted(hp["updated"])
    dbatch = _to_dev(batch)
    tr.zero_grad()
    loss, logits = tr.loss_and_grads(*dbatch)
    assert abs(loss.item() - gold["losses"][0]) <= 2e-5 * abs(gold["losses"][0])
    gu.check(gold, "logits", logits, 2e-5)
    o_loss, _, o_grads = to.head_loss_and_grads(sd, params, *batch, hp["smoothing"], hp["gamma"],
                                                torch.tensor(hp["class_weights"]))
    for name, g in zip(tr.names, tr.grads):
        # fp32 against fp32 with a different summation order; the key bias has a zero gradient (noise both sides)
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 1e-4 * scale + 1e-8, name
        if not name.endswith("in_proj_bias"):
            gu.check(gold, f"grad/{name}", g, 3e-4, what="gradient ")
    # three optimisation steps (the first gradient is already in the buffer)
    losses = [loss.item()]
    tr.step()
    for _ in range(hp["steps"] - 1):
        l, _ = tr.train_step(*dbatch)
        losses.append(l.item())
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 5e-4 * abs(b), (losses, list(gold["losses"]))
    state = fm.state_dict()
    for name in tr.names:
        # Adam divides by sqrt(v): elements whose gradient is rounding noise move by +-lr in either implementation
        check_updated_parameter(gold, name, state[name], 2e-3)
    for name, t in state.items():  # everything off the logits path is untouched
        if name not in tr.names and t.is_floating_point():
            assert torch.equal(t.cpu(), sd[name]), name


@pytest.mark.parametrize("fixture", ["train_head_mask.npz", "train_head_mask_bce.npz"])
def test_mask_dice_term_matches_reference_fixture_and_oracle(fixture):
    """classification + mask terms (train_fusion.py:238-255; "dice" and "dice_bce" mask losses): loss, fused mask
    logits, all 24 gradients and the three-step trajectory against the reference fixture and the oracle."""
    from fusion_train import FusionHeadTrainer

    gold = gu.load(fixture)
    hp = json.loads(str(gold["hp"]))
    params, fm, sd = _head(hp["weight_seed"])
    batch = op.synthetic_head_batch(hp["n"], seed=hp["seed"])
    masks = op.synthetic_raw(hp["n"], seed=hp["seed"] + 1, kind="S")[2]
    tr = FusionHeadTrainer(fm, lr=hp["lr"], betas=hp["betas"], eps=hp["eps"], weight_decay=hp["weight_decay"],
                           smoothing=hp["smoothing"], gamma=hp["gamma"], class_weights=hp["class_weights"],
                           lambda_mask=hp["lambda_mask"], mask_loss_type=hp["mask_loss_type"])
    assert sorted(tr.names) == sorted(hp["updated"]) and len(tr.names) == 24
    dbatch = _to_dev(batch)
    tr.zero_grad()
    loss, logits = tr.loss_and_grads(*dbatch, masks.to(DEV))
    assert abs(loss.item() - gold["losses"][0]) <= 2e-5 * abs(gold["losses"][0])
    gu.check(gold, "logits", logits, 2e-5)
    gu.check(gold, "fused_mask", tr.fused_mask_logits, 1e-4)
    _, _, o_grads = to.head_loss_and_grads(sd, params, *batch, hp["smoothing"], hp["gamma"],
                                           torch.tensor(hp["class_weights"]), masks, hp["lambda_mask"],
                                           hp["mask_loss_type"])
    for name, g in zip(tr.names, tr.grads):
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 2e-4 * scale + 1e-8, name
        if not name.endswith("in_proj_bias"):
            gu.check(gold, f"grad/{name}", g, 4e-4, what="gradient ")
    losses = [loss.item()]
    tr.step()
    for _ in range(hp["steps"] - 1):
        l, _ = tr.train_step(*dbatch, masks.to(DEV))
        losses.append(l.item())
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 5e-4 * abs(b), (losses, list(gold["losses"]))
    state = fm.state_dict()
    for name in tr.names:
        check_updated_parameter(gold, name, state[name], 2e-3)
    # the inference path (bf16 maps, its own mask-head kernels) produces the same fused mask logits
    fm.eval()
    _, mask_inf, _ = fm([dbatch[0]], [dbatch[1]], dbatch[2], dbatch[3])
    tr.zero_grad()
    tr.loss_and_grads(*dbatch, masks.to(DEV))
    assert _rel(mask_inf, tr.fused_mask_logits) < 2e-2


@pytest.mark.parametrize("n", [1, 3])
def test_odd_batch_sizes_match_the_oracle(n):
    """Ragged tiles everywhere (B*16 = 16 / 48 token rows, B rows for the per-case weight gradients), both terms."""
    from fusion_train import FusionHeadTrainer

    params, fm, sd = _head(5)
    batch = op.synthetic_head_batch(n, seed=70 + n)
    masks = op.synthetic_raw(n, seed=80 + n, kind="S")[2]
    tr = FusionHeadTrainer(fm, smoothing=0.1, gamma=2.0, lambda_mask=0.2)
    tr.zero_grad()
    loss, _ = tr.loss_and_grads(*_to_dev(batch), masks.to(DEV))
    o_loss, _, o_grads = to.head_loss_and_grads(sd, params, *batch, 0.1, 2.0, None, masks, 0.2)
    assert abs(loss.item() - float(o_loss)) <= 2e-5 * abs(float(o_loss))
    for name, g in zip(tr.names, tr.grads):
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 2e-4 * scale + 1e-8, name
    empty = tuple(t[:0] for t in _to_dev(batch))
    with pytest.raises(ValueError):
        tr.loss_and_grads(*empty, masks[:0].to(DEV))


def test_vit_geometry_with_overlapping_token_bins_matches_the_oracle():
    """768-channel 14 x 14 maps (the ViT-B/16 encoders, C4): adaptive 4 x 4 pooling has overlapping bins, so GAP(p) is
    not the mean of the tokens - the pooled vectors take their own pass (channel sums) and projection."""
    import model_module as mm
    import parameters_default as pd
    from fusion_train import FusionHeadTrainer

    params = pd.default_parameters()
    fs = params["fusion_model_parameters"]["fusion_specific_parameters"]
    fs["dwi_out_channels"] = fs["dce_out_channels"] = 768
    fm = mm.FusionModel(params)
    sd = op.seeded_state_dict(op.shapes_of(fm.state_dict()), seed=9)
    fm.load_state_dict(sd)
    fm.to(DEV).eval()
    batch = op.synthetic_head_batch(5, seed=55, channels=768, size=14)
    tr = FusionHeadTrainer(fm, smoothing=0.1, gamma=1.5, class_weights=[1.2, 0.8, 1.0, 1.1])
    tr.zero_grad()
    loss, logits = tr.loss_and_grads(*_to_dev(batch))
    o_loss, o_logits, o_grads = to.head_loss_and_grads(sd, params, *batch, 0.1, 1.5, torch.tensor([1.2, 0.8, 1.0, 1.1]))
    assert abs(loss.item() - float(o_loss)) <= 2e-5 * abs(float(o_loss))
    assert _rel(logits.cpu(), o_logits) < 2e-5
    for name, g in zip(tr.names, tr.grads):
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 2e-4 * scale + 1e-8, name
    # mask term on 14 x 14 maps: MaskHeadResize's interpolation dispatch (pre -> bilinear resize to 32 -> out)
    g = torch.Generator().manual_seed(77)
    md, mc = torch.randn(5, 1, 32, 32, generator=g), torch.randn(5, 1, 32, 32, generator=g)
    masks = (torch.rand(5, 1, 32, 32, generator=g) > 0.6).float()
    batch = (batch[0], batch[1], md, mc, batch[4])
    for loss_type in ("dice", "dice_bce"):
        fm.load_state_dict(sd)
        tr = FusionHeadTrainer(fm, smoothing=0.1, gamma=1.5, lambda_mask=0.2, mask_loss_type=loss_type)
        tr.zero_grad()
        loss, _ = tr.loss_and_grads(*_to_dev(batch), masks.to(DEV))
        o_loss, _, o_grads = to.head_loss_and_grads(sd, params, *batch, 0.1, 1.5, None, masks, 0.2, loss_type)
        assert abs(loss.item() - float(o_loss)) <= 2e-5 * abs(float(o_loss)), loss_type
        assert len(tr.names) == 24 and sorted(tr.names) == sorted(o_grads)
        for name, gr in zip(tr.names, tr.grads):
            scale = o_grads[name].abs().max().item()
            assert (gr.cpu() - o_grads[name]).abs().max().item() <= 2e-4 * scale + 1e-8, (loss_type, name)
        fm.eval()
        dbatch = _to_dev(batch)
        _, mask_inf, _ = fm([dbatch[0]], [dbatch[1]], dbatch[2], dbatch[3])
        assert tuple(tr.fused_mask_logits.shape) == (5, 1, 32, 32) and _rel(mask_inf, tr.fused_mask_logits) < 2e-2
    with pytest.raises(NotImplementedError):   # 64-pixel maps go through MaskHeadResize's GELU convolutions
        b64 = op.synthetic_head_batch(2, seed=1, channels=768, size=64)
        FusionHeadTrainer(fm, lambda_mask=0.2).loss_and_grads(*_to_dev(b64), torch.zeros(2, 1, 32, 32, device=DEV))


@pytest.mark.parametrize("cross,se", [(False, True), (True, False), (False, False)])
def test_optional_blocks_off_match_the_oracle(cross, se):
    """FusionModel without its cross-attention block and / or its SE block (use_cross_attention, use_se), both loss
    terms: the trainable set shrinks accordingly and every gradient still matches autograd over the oracle."""
    import model_module as mm
    import parameters_default as pd
    from fusion_train import FusionHeadTrainer

    params = pd.default_parameters()
    params["fusion_model_parameters"]["fusion_specific_parameters"]["use_cross_attention"] = cross
    params["fusion_model_parameters"]["use_se"] = se
    fm = mm.FusionModel(params)
    sd = op.seeded_state_dict(op.shapes_of(fm.state_dict()), seed=13)
    fm.load_state_dict(sd)
    fm.to(DEV).eval()
    batch = op.synthetic_head_batch(6, seed=91)
    masks = op.synthetic_raw(6, seed=92, kind="S")[2]
    tr = FusionHeadTrainer(fm, smoothing=0.1, gamma=1.5, lambda_mask=0.2)
    tr.zero_grad()
    loss, logits = tr.loss_and_grads(*_to_dev(batch), masks.to(DEV))
    o_loss, o_logits, o_grads = to.head_loss_and_grads(sd, params, *batch, 0.1, 1.5, None, masks, 0.2)
    assert sorted(tr.names) == sorted(o_grads)
    assert abs(loss.item() - float(o_loss)) <= 2e-5 * abs(float(o_loss)) and _rel(logits.cpu(), o_logits) < 2e-5
    for name, g in zip(tr.names, tr.grads):
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 2e-4 * scale + 1e-8, name
    first = loss.item()
    tr.lr = 2e-3
    tr.step()
    for _ in range(10):
        last, _ = tr.train_step(*_to_dev(batch), masks.to(DEV))
    assert last.item() < first


def test_larger_batch_gradients_match_oracle_and_training_reduces_the_loss():
    from fusion_train import FusionHeadTrainer

    params, fm, sd = _head(21)
    batch = op.synthetic_head_batch(48, seed=31)
    tr = FusionHeadTrainer(fm, lr=2e-3, weight_decay=0.0, smoothing=0.1, gamma=1.5)
    dbatch = _to_dev(batch)
    tr.zero_grad()
    loss, _ = tr.loss_and_grads(*dbatch)
    o_loss, _, o_grads = to.head_loss_and_grads(sd, params, *batch, 0.1, 1.5, None)
    assert abs(loss.item() - float(o_loss)) <= 2e-5 * abs(float(o_loss))
    for name, g in zip(tr.names, tr.grads):
        scale = o_grads[name].abs().max().item()
        assert (g.cpu() - o_grads[name]).abs().max().item() <= 1e-4 * scale + 1e-8, name
    first = loss.item()
    tr.step()
    for _ in range(15):
        last, _ = tr.train_step(*dbatch)
    assert last.item() < 0.7 * first


def test_lightning_surface_trains_and_inference_sees_the_new_weights():
    import model_module as mm
    import parameters_default as pd
    from train_fusion import LightningFusionModel

    params = pd.default_parameters()
    params["b200_classification_objective_only"] = True
    params["b200_frozen_encoder_mode"] = "eval"   # eval-mode frozen encoders + cls (+ mask) objective: the head trainer
    params["fusion_model_parameters"]["optimizer_parameters"] = {"name": "adamW", "lr": 2e-3, "betas": (0.9, 0.999),
                                                                 "eps": 1e-8, "weight_decay": 4e-5}
    mods = {"dwi": mm.ModelMaskHeadBackbone("dwi", params), "dce": mm.ModelMaskHeadBackbone("dce", params),
            "fusion": mm.FusionModel(params)}
    for k, m in mods.items():
        m.load_state_dict(op.seeded_state_dict(op.shapes_of(m.state_dict()), seed=7))
        m.to(DEV).eval()
    for k in ("dwi", "dce"):
        for p in mods[k].parameters():
            p.requires_grad = False
    import copy
    fresh_head = copy.deepcopy(mods["fusion"])
    lm = LightningFusionModel(mods["dwi"], mods["dce"], mods["fusion"], params)
    dwi_raw, dce, _, labels = op.synthetic_raw(16, seed=3, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    batch = (dwi, dce, labels)
    before = lm.validation_step(batch).item()
    logits0 = lm.forward_from_inputs(dwi.to(DEV), dce.to(DEV))[0].clone()
    losses = [lm.fit_batch(batch).item() for _ in range(12)]
    assert losses[-1] < losses[0]
    # the inference path repacks the updated head weights (bf16 proj_in): its logits follow the trainer's
    _, tr_logits, _, _ = lm._shared_step(batch, "train", return_preds=True)
    logits1 = lm.forward_from_inputs(dwi.to(DEV), dce.to(DEV))[0]
    assert (logits1 - logits0).abs().max().item() > 1e-3
    assert _rel(logits1, tr_logits) < 2e-2
    assert lm.validation_step(batch).item() < before
    # default objective of the reference minus its unbuilt terms: classification + mask dice, 4-tuple batches
    params2 = pd.default_parameters()
    params2["b200_frozen_encoder_mode"] = "eval"
    params2["fusion_model_parameters"]["optimizer_parameters"] = params["fusion_model_parameters"]["optimizer_parameters"]
    lm2 = LightningFusionModel(mods["dwi"], mods["dce"], fresh_head, params2)
    masks = op.synthetic_raw(16, seed=3, kind="S")[2]
    batch4 = (dwi, dce, masks, labels)
    l2 = [lm2.fit_batch(batch4).item() for _ in range(12)]
    assert lm2.head_trainer.lambda_mask == 0.2 and len(lm2.head_trainer.names) == 24
    assert l2[-1] < l2[0]
    # the reference's own configuration: reconstruction + mimic terms on, encoders in train mode (frozen first, then
    # unfrozen through the gradual-unfreeze hook) -> train_graph.FullFusionTrainer behind the same surface
    params3 = pd.default_parameters()
    params3["fusion_model_parameters"].update(recon_enabled=True, lambda_recon=0.1, mimic_enabled=True, lambda_mimic=0.2,
                                              optimizer_parameters=dict(params["fusion_model_parameters"]["optimizer_parameters"],
                                                                        lr=5e-4))
    head3 = copy.deepcopy(fresh_head)
    for m in (mods["dwi"], mods["dce"], head3):
        m.train()
    lm3 = LightningFusionModel(mods["dwi"], mods["dce"], head3, params3)
    l3 = [lm3.fit_batch(batch4).item() for _ in range(8)]
    assert lm3.full_trainer is not None and not lm3.full_trainer.encoders_trainable
    assert all(n.startswith("fusion.") for n in lm3.full_trainer.names) and len(lm3.full_trainer.names) == 35
    assert l3[-1] < l3[0]
    w_before = mods["dwi"].block3.skip[0].weight.detach().clone()
    for m in (mods["dwi"], mods["dce"]):
        for p in m.parameters():
            p.requires_grad = True
    lm3.on_train_epoch_start()                      # gradual unfreeze: the flat buffers are re-bound
    assert lm3.full_trainer.encoders_trainable and len(lm3.full_trainer.names) > 150
    l4 = [lm3.fit_batch(batch4).item() for _ in range(6)]
    assert l4[-1] < l4[0] * 1.05 and not torch.equal(mods["dwi"].block3.skip[0].weight.detach(), w_before)
    for m in (mods["dwi"], mods["dce"]):
        m.eval()


def test_fit_host_equals_step_by_step_training():
    """FusionPipeline.fit_host (pinned host batches, uploads overlapped with the previous step) gives the losses of
    the same steps run one by one on device-resident inputs."""
    import copy

    import model_module as mm
    import parameters_default as pd
    import preprocess_helpers as pre
    from fusion_train import FusionHeadTrainer
    from pipeline import FusionPipeline

    params = pd.default_parameters()
    base = {"dwi": mm.ModelMaskHeadBackbone("dwi", params), "dce": mm.ModelMaskHeadBackbone("dce", params),
            "fusion": mm.FusionModel(params)}
    for m in base.values():
        m.load_state_dict(op.seeded_state_dict(op.shapes_of(m.state_dict()), seed=7))
    nyul = pre.NyulStandardizer()
    nyul.fit(list(op.synthetic_raw(8, seed=6, kind="S")[1]), num_channels=6)
    batches = []
    for i in range(3):
        dwi, dce, _, lab = op.synthetic_raw(8, seed=40 + i, kind="S")
        batches.append((dwi.pin_memory(), dce.pin_memory(), lab.pin_memory()))
    losses = []
    for mode in ("host", "device"):
        mods = {k: copy.deepcopy(m).to(DEV).eval() for k, m in base.items()}
        pipe = FusionPipeline(mods["dwi"], mods["dce"], mods["fusion"], nyul, aux_mode="logits").eval()
        tr = FusionHeadTrainer(mods["fusion"], lr=1e-3)
        if mode == "host":
            losses.append([l.item() for l in pipe.fit_host(batches, tr)])
        else:
            out = []
            for d, c, lab in batches:
                loss, _ = tr.train_step(*pipe.encode_raw(d.to(DEV), c.to(DEV)), lab.to(DEV))
                out.append(loss.item())
            losses.append(out)
    assert len(losses[0]) == 3
    for a, b in zip(*losses):
        assert abs(a - b) <= 1e-4 * abs(b), losses


def test_optimizer_state_roundtrip_and_model_reload():
    """state_dict / load_state_dict of the trainer (AdamW moments + step count) resume training bit for bit on the same
    batch order; load_state_dict on the MODEL writes through the flat-buffer views (the parameters stay bound)."""
    from fusion_train import FusionHeadTrainer

    params, fm, sd = _head(3)
    batch = _to_dev(op.synthetic_head_batch(4, seed=17))
    tr = FusionHeadTrainer(fm, lr=1e-3)
    for _ in range(3):
        tr.train_step(*batch)
    opt_state = tr.state_dict()
    model_state = {k: v.clone() for k, v in fm.state_dict().items()}
    ref = [tr.train_step(*batch)[0].item() for _ in range(2)]
    ref_param = fm.state_dict()["classifier.2.weight"].clone()
    # rewind: reload the model (in place, through the views) and the optimiser state
    fm.load_state_dict(model_state)
    assert tr._bind() is tr._flat and fm.classifier[2].weight.data_ptr() == tr.params[-2].data_ptr()
    tr.load_state_dict(opt_state)
    again = [tr.train_step(*batch)[0].item() for _ in range(2)]
    # split-K weight gradients accumulate with float atomics: equal to fp32 round-off, not bitwise
    assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(again, ref)), (again, ref)
    assert _rel(fm.state_dict()["classifier.2.weight"], ref_param) < 1e-4
    # the dict has torch.optim.AdamW's shape (generic checkpoint code indexes "state" / "param_groups")
    assert set(opt_state) >= {"state", "param_groups"} and opt_state["param_groups"][0]["params"] == list(range(len(tr.params)))
    assert all(set(st) == {"step", "exp_avg", "exp_avg_sq"} for st in opt_state["state"].values())
    assert tuple(opt_state["state"][0]["exp_avg"].shape) == tuple(tr.params[0].shape)
    flat = tr._bind()
    legacy = {"step": 3, "names": list(tr.names), "exp_avg": flat["m"].clone(), "exp_avg_sq": flat["v"].clone()}
    tr.load_state_dict(legacy)                                  # the flat form of earlier checkpoints still loads
    with pytest.raises(ValueError):
        tr.load_state_dict(dict(legacy, names=["x"]))
    with pytest.raises(ValueError):
        tr.load_state_dict(dict(legacy, exp_avg=legacy["exp_avg"][:-8]))
    with pytest.raises(ValueError):
        tr.load_state_dict(dict(opt_state, state={0: opt_state["state"][0]}))
