"""fp32 CPU oracle of the fusion-head fine-tuning step (test infrastructure only - see oracle/__init__.py).

Restates, over a flat state dict, what one optimisation step of the reference does to the fusion head while the
encoders are frozen (`backbone_freeze_on_start`, /root/reference/code/selector_helpers.py:383-386, :496-498) under
the classification objective:

  * FusionModel.forward                      /root/reference/code/model_module.py:919-1000 (oracle.model_oracle)
  * LabelSmoothing                           /root/reference/code/loss.py:190-213
  * SoftFocalLoss / SoftWeightedFocalLoss    /root/reference/code/loss.py:133-188 (reduction "mean")
  * cls_loss = criterion(logits, smoothed)   /root/reference/code/train_fusion.py:238-242
  * mask term: lambda_mask * mean of three SoftDiceLoss   train_fusion.py:245-255, /root/reference/code/loss.py:45-62
  * torch.optim.AdamW (no amsgrad)           /root/reference/code/selector_helpers.py:222-229

Gradients come from torch autograd over the restated full-resolution forward (NOT the pooled-token shortcut the
CUDA path uses), so agreement also checks that shortcut.  Pinned against the unmodified reference modules by
oracle/make_golden_train.py -> tests/golden/train_head.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import model_oracle as mo


def smoothed_targets(labels, num_classes, smoothing):
    """loss.py:203-213: smoothing / (K-1) everywhere, 1 - smoothing on the label."""
    t = torch.full((labels.shape[0], num_classes), smoothing / (num_classes - 1), dtype=torch.float32)
    t.scatter_(1, labels.long().unsqueeze(1), 1.0 - smoothing)
    return t


def soft_focal_loss(logits, targets, gamma, class_weights=None):
    """loss.py:139-152 / :168-186 with reduction "mean"."""
    log_probs = F.log_softmax(logits, dim=1)
    fw = (1 - log_probs.exp()) ** gamma
    if class_weights is not None:
        fw = fw * class_weights.view(1, -1)
    return (-(targets * fw * log_probs).sum(dim=1)).mean()


def soft_dice_loss(logits, targets, eps=1e-6):
    """SoftDiceLoss, loss.py:45-62."""
    probs = torch.sigmoid(logits)
    dims = tuple(range(2, probs.ndim))
    inter = (probs * targets).sum(dims)
    union = probs.sum(dims) + targets.sum(dims)
    return 1.0 - ((2.0 * inter + eps) / (union + eps)).mean()


def dice_bce_loss(logits, targets, eps=1e-6):
    """DiceBCELoss(bce_weight=1, dice_weight=1), loss.py:11-43."""
    bce = F.binary_cross_entropy_with_logits(logits, targets)
    p = torch.sigmoid(logits).flatten(1)
    t = targets.flatten(1)
    dice = (2.0 * (p * t).sum(1)) / (p.sum(1) + t.sum(1) + eps)
    return bce + (1.0 - dice.mean())


def head_loss_and_grads(sd, params, f3_dwi, f3_dce, mask_dwi, mask_dce, labels, smoothing, gamma, class_weights=None,
                        masks=None, lambda_mask=0.0, mask_loss_type="dice"):
    """-> (loss, logits, {name: grad}) for every fusion-head parameter that receives a gradient.  With lambda_mask > 0
    the mask term of train_fusion.py:245-255 is added: lambda_mask * mean of the dice losses of the two encoder masks
    (constants here) and the fused mask."""
    def is_param(k, v):  # BatchNorm running statistics are buffers
        return v.is_floating_point() and k.rsplit(".", 1)[-1] not in ("running_mean", "running_var")

    leaf = {k: (v.detach().clone().requires_grad_(True) if is_param(k, v) else v) for k, v in sd.items()}
    logits, fused_mask, _ = mo.fusion_forward(leaf, params, [f3_dwi], [f3_dce], mask_dwi, mask_dce)
    targets = smoothed_targets(labels, logits.shape[1], smoothing)
    loss = soft_focal_loss(logits, targets, gamma, class_weights)
    if lambda_mask > 0:
        crit = soft_dice_loss if mask_loss_type == "dice" else dice_bce_loss  # selector_helpers.py:103-106
        loss = loss + lambda_mask * (crit(mask_dwi, masks) + crit(mask_dce, masks) + crit(fused_mask, masks)) / 3
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if is_param(k, v) and v.grad is not None}
    return loss.detach(), logits.detach(), grads


def adamw_step(p, g, m, v, step, lr, betas, eps, weight_decay):
    """torch.optim.AdamW single-tensor update, restated; returns (p, m, v) new tensors."""
    b1, b2 = betas
    p = p * (1 - lr * weight_decay)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def train_steps(sd, params, batch, steps, smoothing, gamma, class_weights, lr, betas, eps, weight_decay, masks=None,
                lambda_mask=0.0, mask_loss_type="dice"):
    """`steps` AdamW steps on one batch.  -> (losses, final state dict, names that were updated)."""
    sd = {k: v.clone() for k, v in sd.items()}
    state = {}
    losses, names = [], []
    for it in range(1, steps + 1):
        loss, _, grads = head_loss_and_grads(sd, params, *batch, smoothing, gamma, class_weights, masks, lambda_mask,
                                             mask_loss_type)
        losses.append(float(loss))
        names = sorted(grads)
        for k, g in grads.items():
            m, v = state.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
            sd[k], m, v = adamw_step(sd[k], g, m, v, it, lr, betas, eps, weight_decay)
            state[k] = (m, v)
    return losses, sd, names
