"""Generates tests/golden/train_head.npz from the UNMODIFIED reference (authoring container only):

    python oracle/make_golden_train.py      # needs /root/reference (read-only)

The reference's FusionModel (code/model_module.py), LabelSmoothing + SoftWeightedFocalLoss (code/loss.py) and
torch.optim.AdamW - the pieces LightningFusionModel._shared_step / configure_optimizers (code/train_fusion.py:203-242,
code/selector_helpers.py:222-229) put together for the always-trainable fusion-head group - run on seeded inputs and
seeded weights; loss, logits, every parameter gradient and the parameters after three optimisation steps are stored
(full tensors when small, strided probes + sums otherwise).  `main_full` additionally runs the reference's whole
frozen-phase objective - classification + mask dice + fused reconstruction + mimic, with the loss functions imported from
the unmodified code/train_fusion.py and code/train.py behind import stubs for the absent harness packages - and stores
the total, its terms and all 35 gradients (tests/golden/train_head_full.npz): the CUDA path builds the first two terms,
the fixture pins the oracle for the other two ahead of them.  Nothing here is imported by the product.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import make_golden as mg  # noqa: E402  (also puts /root/reference/code on sys.path)
from oracle import params as op  # noqa: E402

HP = {"n": 8, "seed": 11, "smoothing": 0.1, "gamma": 1.5, "class_weights": [0.7, 1.3, 1.0, 0.9], "lr": 1e-3,
      "betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 1e-2, "steps": 3, "weight_seed": 7}


def main(lambda_mask=0.0, out_name="train_head.npz", mask_loss_type="dice"):
    import loss as ref_loss
    import model_module as mm

    p = mg.configure(mg.reference_parameters())
    torch.manual_seed(0)
    model = mm.FusionModel(p)
    model.load_state_dict(op.seeded_state_dict(op.shapes_of(model.state_dict()), seed=HP["weight_seed"]))
    model.train()  # what Lightning does; the logits path holds no BatchNorm / Dropout
    f3d, f3c, md, mc, labels = op.synthetic_head_batch(HP["n"], seed=HP["seed"])
    smoother = ref_loss.LabelSmoothing(p["class_num"], HP["smoothing"])
    crit = ref_loss.SoftWeightedFocalLoss(HP["gamma"], torch.tensor(HP["class_weights"]))
    opt = torch.optim.AdamW(model.parameters(), lr=HP["lr"], betas=tuple(HP["betas"]), eps=HP["eps"],
                            weight_decay=HP["weight_decay"], amsgrad=False)
    masks = op.synthetic_raw(HP["n"], seed=HP["seed"] + 1, kind="S")[2]  # [n,1,32,32] in {0,1}
    dice = ref_loss.SoftDiceLoss() if mask_loss_type == "dice" else ref_loss.DiceBCELoss(bce_weight=1.0, dice_weight=1.0)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    out, losses = {}, []
    for it in range(HP["steps"]):
        opt.zero_grad(set_to_none=True)
        logits, fused_mask, _ = model([f3d], [f3c], md, mc)
        loss = crit(logits, smoother(logits, labels))
        if lambda_mask > 0:  # train_fusion.py:245-255 (safe_mask_loss = the criterion at equal sizes)
            loss = loss + lambda_mask * (dice(md, masks) + dice(mc, masks) + dice(fused_mask, masks)) / 3
        loss.backward()
        losses.append(float(loss))
        if it == 0:
            mg.flatten("logits", logits, out)
            mg.flatten("fused_mask", fused_mask, out)
            for k, v in model.named_parameters():
                if v.grad is not None:
                    mg.flatten(f"grad/{k}", v.grad, out)
        opt.step()
    updated = [k for k, v in model.named_parameters() if not torch.equal(v.detach(), before[k])]
    for k, v in model.named_parameters():
        if k in updated:
            mg.flatten(f"param/{k}", v, out)
    out["losses"] = np.array(losses, dtype=np.float64)
    out["hp"] = np.array(json.dumps(dict(HP, updated=updated, lambda_mask=lambda_mask, mask_loss_type=mask_loss_type)))
    np.savez_compressed(os.path.join(mg.GOLD, out_name), **out)
    print("losses", losses)
    print("updated", len(updated), "parameters:", updated)


def stub_harness_modules():
    """pytorch_lightning / torchmetrics / matplotlib are absent here; train.py and train_fusion.py only need the names
    at import time.  The stubs carry no arithmetic."""
    import types

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return None

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Any()

    class LightningModule(torch.nn.Module):
        pass

    def anything(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any

    def any_instance(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()

    pl = mod("pytorch_lightning", LightningModule=LightningModule, Trainer=_Any, seed_everything=lambda *a, **k: None)
    pl.callbacks = mod("pytorch_lightning.callbacks", ModelCheckpoint=_Any, LearningRateMonitor=_Any,
                       EarlyStopping=_Any, Callback=_Any)
    pl.loggers = mod("pytorch_lightning.loggers", TensorBoardLogger=_Any, CSVLogger=_Any)
    tm = mod("torchmetrics", MeanMetric=_Any, Metric=_Any)
    tm.__getattr__ = anything
    cl = mod("torchmetrics.classification")
    cl.__getattr__ = anything
    sg = mod("torchmetrics.segmentation")
    sg.__getattr__ = anything
    mp = mod("matplotlib")
    mp.pyplot = mod("matplotlib.pyplot")
    mp.pyplot.__getattr__ = any_instance


def main_full(out_name="train_head_full.npz"):
    """FusionModel in train mode + the reference's OWN loss functions (imported from code/train_fusion.py and
    code/train.py behind import stubs): classification + mask dice + fused reconstruction + mimic, one backward."""
    stub_harness_modules()
    import loss as ref_loss
    import model_module as mm
    import train as ref_train
    import train_fusion as ref_tf

    lam = {"lambda_mask": 0.2, "lambda_recon": 0.1, "lambda_mimic": 0.2}   # parameters_generate.py:108-125
    p = mg.configure(mg.reference_parameters())
    torch.manual_seed(0)
    model = mm.FusionModel(p)
    model.load_state_dict(op.seeded_state_dict(op.shapes_of(model.state_dict()), seed=HP["weight_seed"]))
    model.train()
    f3d, f3c, md, mc, labels = op.synthetic_head_batch(HP["n"], seed=HP["seed"])
    dwi_in, dce_in, masks, _ = op.synthetic_raw(HP["n"], seed=HP["seed"] + 1, kind="S")
    dwi_in = dwi_in / dwi_in.amax(dim=(1, 2, 3), keepdim=True)
    smoother = ref_loss.LabelSmoothing(p["class_num"], HP["smoothing"])
    crit = ref_loss.SoftWeightedFocalLoss(HP["gamma"], torch.tensor(HP["class_weights"]))
    dice = ref_loss.SoftDiceLoss()
    logits, fused_mask, aux = model([f3d], [f3c], md, mc)
    cls = crit(logits, smoother(logits, labels))
    mask = (ref_tf.safe_mask_loss(md, masks, dice) + ref_tf.safe_mask_loss(mc, masks, dice) +
            ref_tf.safe_mask_loss(fused_mask, masks, dice)) / 3
    fused_input = torch.cat([dwi_in, dce_in], dim=1)
    recon = (ref_tf.compute_recon_list_loss(None, dwi_in) + ref_tf.compute_recon_list_loss(None, dce_in) +
             ref_tf.compute_recon_list_loss(aux["recon_fused"], fused_input)) / 3
    p1, p1_r, p2, p2_r = aux["proj_fused"][:4]                                  # train_fusion.py:287-290, as written
    mimic = (ref_train.mimic_feat_loss(p1, p1_r) + ref_train.mimic_feat_loss(p2, p2_r)) / 2
    total = cls + lam["lambda_mask"] * mask + lam["lambda_recon"] * recon + lam["lambda_mimic"] * mimic
    total.backward()
    out = {}
    mg.flatten("logits", logits, out)
    mg.flatten("recon_fused", aux["recon_fused"], out)
    mg.flatten("proj_fused", aux["proj_fused"], out)
    names = []
    for k, v in model.named_parameters():
        if v.grad is not None:
            mg.flatten(f"grad/{k}", v.grad, out)
            names.append(k)
    out["parts"] = np.array([t.item() for t in (total, cls, mask, recon, mimic)], dtype=np.float64)
    out["hp"] = np.array(json.dumps(dict(HP, **lam, with_grad=names)))
    np.savez_compressed(os.path.join(mg.GOLD, out_name), **out)
    print("full objective", out["parts"], len(names), "tensors with a gradient")


def main_c1(out_name="train_c1_dwi.npz", n=8):
    """BASELINE config C1: the DWI CNN encoder's forward + training loss + backward on the CPU, by the reference's
    own modules and loss functions (code/train.py behind the import stubs), dropouts at p = 0 so that the train-mode
    forward is deterministic.  Composed exactly as LightningSingleModel._shared_step does (train.py:294-400)."""
    import types

    stub_harness_modules()
    import loss as ref_loss
    import model_module as mm
    import train as ref_train

    p = mg.configure(mg.reference_parameters())
    mp = p["dwi_model_parameters"]
    mp["dropout"] = 0.0
    lam = {"lambda_mask": mp["mask_parameters"]["lambda_mask"], "lambda_recon": mp["lambda_recon"],
           "lambda_mimic": mp["lambda_mimic"], "lambda_feat_norm": mp["lambda_feat_norm"]}
    torch.manual_seed(0)
    model = mm.ModelMaskHeadBackbone("dwi", p, None)
    model.load_state_dict(op.seeded_state_dict(op.shapes_of(model.state_dict()), seed=HP["weight_seed"]))
    model.train()
    dwi_raw, _, masks, labels = op.synthetic_raw(n, seed=1234, kind="S")
    x = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    outputs, aux, mask_output = model(x, masks)
    smoother = ref_loss.LabelSmoothing(p["class_num"], HP["smoothing"])
    crit = ref_loss.SoftWeightedFocalLoss(HP["gamma"], torch.tensor(HP["class_weights"]))
    cls = crit(outputs, smoother(outputs, labels))
    feat_norm = ref_train.compute_feat_norm_loss(aux, x.device)
    mask = ref_loss.SoftDiceLoss()(mask_output, masks)
    ns = types.SimpleNamespace(device=x.device, mimic_enabled=True, lambda_recon=lam["lambda_recon"],
                               lambda_mimic=lam["lambda_mimic"])
    recon_w, mimic_w = ref_train.LightningSingleModel.compute_aux_losses(ns, aux, x, aux["proj_pairs"], 1.0, True)
    total = (cls + lam["lambda_feat_norm"] * feat_norm + lam["lambda_mask"] * mask +
             lam["lambda_recon"] * recon_w * 1.0 + lam["lambda_mimic"] * mimic_w * 1.0)       # train.py:396-399
    total.backward()
    out, names = {}, []
    mg.flatten("logits", outputs, out)
    mg.flatten("mask", mask_output, out)
    for k, v in model.named_parameters():
        if v.grad is not None:
            mg.flatten(f"grad/{k}", v.grad, out)
            names.append(k)
    out["parts"] = np.array([t.item() for t in (total, cls, feat_norm, mask, recon_w, mimic_w)], dtype=np.float64)
    out["hp"] = np.array(json.dumps(dict(HP, **lam, n=n, with_grad=names)))
    np.savez_compressed(os.path.join(mg.GOLD, out_name), **out)
    print("C1 objective", out["parts"], len(names), "tensors with a gradient")


if __name__ == "__main__":
    main()
    main(lambda_mask=0.2, out_name="train_head_mask.npz")
    main(lambda_mask=0.2, out_name="train_head_mask_bce.npz", mask_loss_type="dice_bce")
    main_full()
    main_c1()
