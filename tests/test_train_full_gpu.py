"""GPU: the reference's WHOLE fusion objective (classification + three dice terms + three reconstruction terms + the
mimic term, code/train_fusion.py:238-296) on the training kernels:

* frozen-encoder phase: FusionModel in train mode (batch-statistic BatchNorm in the reconstruction head and the
  projector) against tests/golden/train_head_full.npz, which the reference's own modules and loss functions produced
  (total, every term, all 35 gradients);
* everything unfrozen (BASELINE config C5): both encoders + the head, one step, against the CPU oracle composed from
  the pinned pieces (oracle.model_oracle in batch-statistic mode + oracle.train_oracle's loss restatements).
"""
import json

import numpy as np
import pytest
import torch

import b200path  # noqa: F401
import golden_util as gu
from oracle import model_oracle as mo
from oracle import params as op
from oracle import train_oracle as to

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() / max(b.double().abs().max().item(), 1e-12)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)


def test_frozen_phase_full_objective_vs_reference_fixture():
    import model_module as mm
    import parameters_default as pd
    import train_graph as tg

    gold = gu.load("train_head_full.npz")
    hp = json.loads(str(gold["hp"]))
    p = pd.default_parameters()
    fm = mm.FusionModel(p)
    sd = op.seeded_state_dict(op.shapes_of(fm.state_dict()), seed=hp["weight_seed"])
    fm.load_state_dict(sd)
    fm.to(DEV).train()
    f3d, f3c, md, mc, labels = op.synthetic_head_batch(hp["n"], seed=hp["seed"])
    dwi_in, dce_in, masks, _ = op.synthetic_raw(hp["n"], seed=hp["seed"] + 1, kind="S")
    dwi_in = dwi_in / dwi_in.amax(dim=(1, 2, 3), keepdim=True)
    tr = tg.FullFusionTrainer(None, None, fm, smoothing=hp["smoothing"], gamma=hp["gamma"],
                              class_weights=torch.tensor(hp["class_weights"]), lambda_mask=hp["lambda_mask"],
                              lambda_recon=hp["lambda_recon"], lambda_mimic=hp["lambda_mimic"], encoders_trainable=False)
    tr.zero_grad()
    total, parts = tr.forward_backward(dwi_in.to(DEV), dce_in.to(DEV), masks, labels, md=md[:, 0].contiguous().to(DEV),
                                       mc=mc[:, 0].contiguous().to(DEV), f3d=_nhwc(f3d), f3c=_nhwc(f3c))
    torch.cuda.synchronize()
    g_total, g_cls, g_mask, g_recon, g_mimic = gold["parts"]
    got = {k: v.item() for k, v in parts.items()}
    print("parts (gpu):", got, "reference:", dict(cls=g_cls, mask=g_mask, recon=g_recon, mimic=g_mimic), "total", total.item(), g_total)
    assert abs(total.item() - g_total) <= 5e-3 * abs(g_total)
    for k, want in (("cls", g_cls), ("mask", g_mask), ("recon", g_recon), ("mimic", g_mimic)):
        assert abs(got[k] - want) <= 1e-2 * max(abs(want), 1e-3), (k, got[k], want)
    named = {f"fusion.{k}": v for k, v in fm.named_parameters()}
    errs = {}
    for k in hp["with_grad"]:
        g = named[f"fusion.{k}"].grad
        assert g is not None and f"fusion.{k}" in tr.names, k
        # bf16 maps / bf16 map gradients against the fp32 reference.  The cross-attention block's small tensors get
        # their gradient only through the 4 x 4-token pooling of a bf16 gradient map (column sums over 128 token rows
        # of strongly cancelling terms; the key third of in_proj_bias is mathematically zero): noise-level bound 4e-1; every
        # convolution / BatchNorm / SE / gating / classifier gradient: 6e-2.
        errs[k] = gu.check(gold, f"grad/{k}", g, rtol=4e-1 if k.startswith("cross_attn_block.") else 6e-2)
    assert len(errs) == 35
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print("worst gradient errors:", [(k, f"{v:.2e}") for k, v in worst[:8]])
    assert np.median(list(errs.values())) <= 1.5e-2
    # parameters outside the objective's reach were left out of the flat buffers
    assert not any(n.startswith(("fusion.refine.", "fusion.fusion_conv_reduce.")) for n in tr.names)
    # one AdamW step moves exactly the trainable set
    before = {k: v.detach().clone() for k, v in fm.named_parameters()}
    tr.step()
    moved = [k for k, v in fm.named_parameters() if not torch.equal(v.detach(), before[k])]
    assert sorted(moved) == sorted(hp["with_grad"])


def _oracle_full_step(sds, p, dwi, dce, masks, labels, hp):
    """Autograd over the oracle's batch-statistic forward of both encoders and the head + the reference's objective."""
    def is_param(k, v):
        return v.is_floating_point() and k.rsplit(".", 1)[-1] not in ("running_mean", "running_var")

    leaf = {m: {k: (v.detach().clone().requires_grad_(True) if is_param(k, v) else v) for k, v in sd.items()}
            for m, sd in sds.items()}
    mo.BN_BATCH_STATS = True
    try:
        _, ad, md = mo.encoder_forward(leaf["dwi"], "dwi", p, dwi)
        _, ac, mc = mo.encoder_forward(leaf["dce"], "dce", p, dce)
        logits, fused_mask, aux = mo.fusion_forward(leaf["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)
    finally:
        mo.BN_BATCH_STATS = False
    cw = torch.tensor(hp["class_weights"])
    cls = to.soft_focal_loss(logits, to.smoothed_targets(labels, logits.shape[1], hp["smoothing"]), hp["gamma"], cw)
    mask = (to.soft_dice_loss(md, masks) + to.soft_dice_loss(mc, masks) + to.soft_dice_loss(fused_mask, masks)) / 3

    def rlist(rs, x):
        return sum(to.recon_list_loss(r, x) for r in rs) / len(rs)

    recon = (rlist(ad["recon_feats"], dwi) + rlist(ac["recon_feats"], dce) +
             to.recon_list_loss(aux["recon_fused"], torch.cat([dwi, dce], dim=1))) / 3
    p1, p1_r, p2, p2_r = aux["proj_fused"][:4]
    mimic = (to.mimic_feat_loss(p1, p1_r) + to.mimic_feat_loss(p2, p2_r)) / 2
    total = cls + hp["lambda_mask"] * mask + hp["lambda_recon"] * recon + hp["lambda_mimic"] * mimic
    total.backward()
    grads = {f"{m}.{k}": v.grad for m, d in leaf.items() for k, v in d.items() if is_param(k, v) and v.grad is not None}
    return total.item(), dict(cls=cls.item(), mask=mask.item(), recon=recon.item(), mimic=mimic.item()), grads


def test_c5_unfrozen_step_vs_oracle():
    import model_module as mm
    import parameters_default as pd
    import train_graph as tg

    hp = {"smoothing": 0.1, "gamma": 1.5, "class_weights": [0.7, 1.3, 1.0, 0.9], "lambda_mask": 0.2, "lambda_recon": 0.1,
          "lambda_mimic": 0.2}
    p = pd.default_parameters()
    for m in ("dwi", "dce", "fusion"):
        p[f"{m}_model_parameters"]["dropout"] = 0.0   # deterministic train-mode forward (as the C1 fixture)
    mods = {"dwi": mm.ModelMaskHeadBackbone("dwi", p), "dce": mm.ModelMaskHeadBackbone("dce", p), "fusion": mm.FusionModel(p)}
    sds = {}
    for k, m in mods.items():
        sds[k] = op.seeded_state_dict(op.shapes_of(m.state_dict()), seed=7)
        m.load_state_dict(sds[k])
        m.to(DEV).train()
    n = 6
    dwi_raw, dce, masks, labels = op.synthetic_raw(n, seed=99, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    tr = tg.FullFusionTrainer(mods["dwi"], mods["dce"], mods["fusion"], smoothing=hp["smoothing"], gamma=hp["gamma"],
                              class_weights=torch.tensor(hp["class_weights"]), lambda_mask=hp["lambda_mask"],
                              lambda_recon=hp["lambda_recon"], lambda_mimic=hp["lambda_mimic"])
    tr.zero_grad()
    total, parts = tr.forward_backward(dwi.to(DEV), dce.to(DEV), masks, labels)
    torch.cuda.synchronize()
    o_total, o_parts, o_grads = _oracle_full_step(sds, p, dwi, dce, masks, labels, hp)
    got = {k: v.item() for k, v in parts.items()}
    print("parts (gpu):", got, "oracle:", o_parts, "total", total.item(), o_total)
    assert abs(total.item() - o_total) <= 1e-2 * abs(o_total)
    for k in o_parts:
        assert abs(got[k] - o_parts[k]) <= 2e-2 * max(abs(o_parts[k]), 1e-3), (k, got[k], o_parts[k])
    named = dict(zip(tr.names, tr.params))
    missing = [k for k in o_grads if k not in named and o_grads[k].abs().max().item() > 0]
    assert not missing, missing   # every tensor the reference's autograd reaches is in the trainable set ...
    extra = [k for k in named if k not in o_grads]
    assert not extra, extra       # ... and nothing else
    errs = {k: _rel(named[k].grad, o_grads[k]) for k in o_grads if k in named}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print(len(errs), "gradients; worst:", [(k, f"{v:.2e}") for k, v in worst[:12]])
    noisy = {k: v for k, v in errs.items() if ".mask_spatial_attention." in k}
    rest = {k: v for k, v in errs.items() if k not in noisy}
    assert np.median(list(errs.values())) <= 2e-2
    assert max(rest.values()) <= 8e-2, worst[:6]
    # The handful of tiny mask-attention parameters (a scalar gamma, 16-vectors) are sums over every pixel and channel of
    # products of bf16 gradient maps that cancel almost completely, accumulated with float atomics: their error is the
    # rounding noise of those maps and moves run to run (observed 1-32 % of the tensor's max over repeated runs).  They
    # are sanity-checked - right sign and size - not pinned: the worst below 50 %, their median below 15 %.
    assert max(noisy.values()) <= 5e-1 and np.median(list(noisy.values())) <= 1.5e-1, noisy
    # one optimisation step changes every trainable tensor and nothing else
    before = {k: v.detach().clone() for k, v in named.items()}
    tr.step()
    torch.cuda.synchronize()
    assert all(not torch.equal(named[k].detach(), before[k]) for k in named)
    for tag, m in mods.items():
        for k, v in m.named_parameters():
            if f"{tag}.{k}" not in named:
                assert torch.equal(v.detach().cpu(), sds[tag][k]), k


def test_reference_style_autograd_loop_drives_the_training_kernels():
    """The drop-in claim for training: code written against the reference modules - `model.train()`, forward,
    a torch loss on the outputs, `loss.backward()` (the shape of train.py / train_fusion.py's `_shared_step`) - runs the
    explicit backward pass through the modules' autograd bridge.  Same objective as test_c5_unfrozen_step_vs_oracle,
    losses written with torch operators on the module outputs; gradients against the CPU oracle."""
    import model_module as mm
    import parameters_default as pd

    hp = {"smoothing": 0.1, "gamma": 1.5, "class_weights": [0.7, 1.3, 1.0, 0.9], "lambda_mask": 0.2, "lambda_recon": 0.1,
          "lambda_mimic": 0.2}
    p = pd.default_parameters()
    for m in ("dwi", "dce", "fusion"):
        p[f"{m}_model_parameters"]["dropout"] = 0.0
    mods = {"dwi": mm.ModelMaskHeadBackbone("dwi", p), "dce": mm.ModelMaskHeadBackbone("dce", p), "fusion": mm.FusionModel(p)}
    sds = {}
    for k, m in mods.items():
        sds[k] = op.seeded_state_dict(op.shapes_of(m.state_dict()), seed=7)
        m.load_state_dict(sds[k])
        m.to(DEV).train()
    n = 6
    dwi_raw, dce, masks, labels = op.synthetic_raw(n, seed=99, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    x_d, x_c, mk, lb = dwi.to(DEV), dce.to(DEV), masks.to(DEV), labels.to(DEV)
    # ---- what LightningFusionModel._shared_step does (train_fusion.py:226-296), with torch operators ----
    _, dwi_aux, dwi_mask = mods["dwi"](x_d)
    _, dce_aux, dce_mask = mods["dce"](x_c)
    assert dwi_aux["raw_feats"][2].requires_grad and dwi_mask.requires_grad and dwi_aux["proj_pairs"][0].shape[-1] == 64
    logits, fused_mask, aux = mods["fusion"](dwi_aux["raw_feats"], dce_aux["raw_feats"], dwi_mask, dce_mask)
    cw = torch.tensor(hp["class_weights"], device=DEV)
    tgt = to.smoothed_targets(labels, 4, hp["smoothing"]).to(DEV)
    cls = to.soft_focal_loss(logits, tgt, hp["gamma"], cw)
    mask = (to.soft_dice_loss(dwi_mask, mk) + to.soft_dice_loss(dce_mask, mk) + to.soft_dice_loss(fused_mask, mk)) / 3

    def rlist(rs, x):
        return sum(to.recon_list_loss(r, x) for r in rs) / len(rs)

    recon = (rlist(dwi_aux["recon_feats"], x_d) + rlist(dce_aux["recon_feats"], x_c) +
             to.recon_list_loss(aux["recon_fused"], torch.cat([x_d, x_c], dim=1))) / 3
    p1, p1_r, p2, p2_r = aux["proj_fused"][:4]
    mimic = (to.mimic_feat_loss(p1.float(), p1_r.float()) + to.mimic_feat_loss(p2.float(), p2_r.float())) / 2
    total = cls + hp["lambda_mask"] * mask + hp["lambda_recon"] * recon + hp["lambda_mimic"] * mimic
    total.backward()
    torch.cuda.synchronize()
    o_total, o_parts, o_grads = _oracle_full_step(sds, p, dwi, dce, masks, labels, hp)
    assert abs(total.item() - o_total) <= 1e-2 * abs(o_total), (total.item(), o_total)
    named = {f"{tag}.{k}": v for tag, m in mods.items() for k, v in m.named_parameters()}
    errs = {k: _rel(named[k].grad, g) for k, g in o_grads.items() if named[k].grad is not None}
    missing = [k for k, g in o_grads.items() if named[k].grad is None and g.abs().max().item() > 0]
    assert not missing, missing
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print(len(errs), "gradients through the autograd bridge; worst:", [(k, f"{v:.2e}") for k, v in worst[:8]])
    noisy = {k: v for k, v in errs.items() if ".mask_spatial_attention." in k or ".cross_attn_block." in k}
    rest = {k: v for k, v in errs.items() if k not in noisy}
    assert np.median(list(errs.values())) <= 2e-2 and max(rest.values()) <= 8e-2, worst[:6]
    assert max(noisy.values()) <= 5e-1 and np.median(list(noisy.values())) <= 1.5e-1, noisy   # (see the comment above)
