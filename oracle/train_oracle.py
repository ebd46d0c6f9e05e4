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


# ----------------------------------------------------------------------------------------------------------------
# the rest of the reference's frozen-phase objective (NOT built in the CUDA path yet: this pins the oracle ahead of it)
# ----------------------------------------------------------------------------------------------------------------
def charbonnier_loss(pred, target, eps=1e-3):
    """train.py:1041-1042."""
    return torch.mean(torch.sqrt((pred - target) ** 2 + eps ** 2))


def recon_list_loss(recon, input_img):
    """compute_recon_list_loss for one reconstruction tensor (train_fusion.py:709-745) with recon_image_loss
    (train.py:1043-1048): bilinear up-sample to the input size, channel means when the channel counts differ,
    sigmoid -> clamp -> Charbonnier."""
    if recon is None:
        return torch.zeros(())
    r_up = F.interpolate(recon, size=input_img.shape[-2:], mode="bilinear", align_corners=False)
    target = input_img
    if r_up.size(1) != input_img.size(1):
        r_up, target = r_up.mean(dim=1, keepdim=True), input_img.mean(dim=1, keepdim=True)
    return charbonnier_loss(torch.sigmoid(r_up).clamp(0, 1), target.clamp(0, 1))


def mimic_feat_loss(s_feat, t_feat, eps=1e-6):
    """train.py:1033-1038 (teacher detached)."""
    s = F.normalize(s_feat.flatten(1), dim=1)
    t = F.normalize(t_feat.detach().flatten(1), dim=1)
    return (1.0 - (s * t).sum(dim=1).clamp(-1 + eps, 1 - eps)).mean()


def full_objective_and_grads(sd, params, f3_dwi, f3_dce, mask_dwi, mask_dce, labels, masks, dwi_in, dce_in, smoothing,
                             gamma, class_weights, lambda_mask, lambda_recon, lambda_mimic, aux_w=1.0):
    """Total training loss of LightningFusionModel._shared_step (train_fusion.py:238-292) as far as it depends on the
    fusion head, FusionModel in TRAIN mode (batch-statistic BatchNorm in the reconstruction head and the projector):
    classification + lambda_mask * mean of three dice terms + lambda_recon * recon_fused / 3 + lambda_mimic * mimic
    (the encoders' own reconstruction lists are taken as absent, which compute_recon_list_loss maps to 0).  The
    reference's mimic term unpacks `proj_fused[:4]` - the FIRST FOUR CASES of the fused projection - as
    (p1, p1_r, p2, p2_r) (train_fusion.py:287-290): reproduced as is.  -> (loss, parts, {name: grad})."""
    def is_param(k, v):
        return v.is_floating_point() and k.rsplit(".", 1)[-1] not in ("running_mean", "running_var")

    leaf = {k: (v.detach().clone().requires_grad_(True) if is_param(k, v) else v) for k, v in sd.items()}
    mo.BN_BATCH_STATS = True
    try:
        logits, fused_mask, aux = mo.fusion_forward(leaf, params, [f3_dwi], [f3_dce], mask_dwi, mask_dce)
    finally:
        mo.BN_BATCH_STATS = False
    cls = soft_focal_loss(logits, smoothed_targets(labels, logits.shape[1], smoothing), gamma, class_weights)
    mask = (soft_dice_loss(mask_dwi, masks) + soft_dice_loss(mask_dce, masks) + soft_dice_loss(fused_mask, masks)) / 3
    recon = recon_list_loss(aux["recon_fused"], torch.cat([dwi_in, dce_in], dim=1)) / 3
    p1, p1_r, p2, p2_r = aux["proj_fused"][:4]
    mimic = (mimic_feat_loss(p1, p1_r) + mimic_feat_loss(p2, p2_r)) / 2
    loss = cls + lambda_mask * mask + lambda_recon * recon * aux_w + lambda_mimic * mimic * aux_w
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if is_param(k, v) and v.grad is not None}
    parts = {"cls": cls.item(), "mask": mask.item(), "recon": recon.item(), "mimic": mimic.item()}
    return loss.detach(), parts, grads


def single_model_objective_and_grads(sd, params, method, x, masks, labels, smoothing, gamma, class_weights, lambda_mask,
                                     lambda_recon, lambda_mimic, lambda_feat_norm, aux_w=1.0):
    """Training loss of LightningSingleModel._shared_step (train.py:294-400, :434-466) for one modality's encoder in
    TRAIN mode with its dropouts at p = 0 (BASELINE config C1: the DWI CNN forward + train step on the CPU):
    classification + lambda_feat_norm * sum_f mean(f^2) + lambda_mask * dice(mask logits) + recon + mimic, where the
    reference weights the last two TWICE - compute_aux_losses already multiplies by lambda * aux_w (:462-464) and
    _shared_step multiplies again (:396-399): reproduced as is.  -> (loss, parts, {name: grad})."""
    def is_param(k, v):
        return v.is_floating_point() and k.rsplit(".", 1)[-1] not in ("running_mean", "running_var")

    leaf = {k: (v.detach().clone().requires_grad_(True) if is_param(k, v) else v) for k, v in sd.items()}
    mo.BN_BATCH_STATS = True
    try:
        logits, aux, mask_pred = mo.encoder_forward(leaf, method, params, x)
    finally:
        mo.BN_BATCH_STATS = False
    cls = soft_focal_loss(logits, smoothed_targets(labels, logits.shape[1], smoothing), gamma, class_weights)
    feat_norm = sum(f.pow(2).mean() for f in aux["raw_feats"])                            # train.py:1021-1030
    mask = soft_dice_loss(mask_pred, masks)                                               # train.py:373-374
    recon = torch.zeros(())
    for r in aux["recon_feats"]:                                                          # train.py:446-454
        if r is None:
            continue
        target = x
        r_up = F.interpolate(r, size=target.shape[-2:], mode="bilinear", align_corners=False)
        if r_up.size(1) == 1 and target.size(1) > 1:
            target = target.mean(dim=1, keepdim=True)
        recon = recon + charbonnier_loss(torch.sigmoid(r_up).clamp(0, 1), target.clamp(0, 1))
    p1, p1_r, p2, p2_r = aux["proj_pairs"][:4]
    mimic = mimic_feat_loss(p1, p1_r) + mimic_feat_loss(p2, p2_r)                         # train.py:457-459
    loss = (cls + lambda_feat_norm * feat_norm + lambda_mask * mask +
            lambda_recon * (recon * lambda_recon * aux_w) * aux_w + lambda_mimic * (mimic * lambda_mimic * aux_w) * aux_w)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if is_param(k, v) and v.grad is not None}
    parts = {"cls": cls.item(), "feat_norm": feat_norm.item(), "mask": mask.item(), "recon": recon.item(),
             "mimic": mimic.item()}
    return loss.detach(), parts, grads
