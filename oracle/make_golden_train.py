"""Generates tests/golden/train_head.npz from the UNMODIFIED reference (authoring container only):

    python oracle/make_golden_train.py      # needs /root/reference (read-only)

The reference's FusionModel (code/model_module.py), LabelSmoothing + SoftWeightedFocalLoss (code/loss.py) and
torch.optim.AdamW - the pieces LightningFusionModel._shared_step / configure_optimizers (code/train_fusion.py:203-242,
code/selector_helpers.py:222-229) put together for the always-trainable fusion-head group - run on seeded inputs and
seeded weights; loss, logits, every parameter gradient and the parameters after three optimisation steps are stored
(full tensors when small, strided probes + sums otherwise).  Nothing here is imported by the product.
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


if __name__ == "__main__":
    main()
    main(lambda_mask=0.2, out_name="train_head_mask.npz")
    main(lambda_mask=0.2, out_name="train_head_mask_bce.npz", mask_loss_type="dice_bce")
