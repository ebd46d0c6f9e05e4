"""Inference-time surface of the reference's `LightningFusionModel` (code/train_fusion.py): the three models behind
one object, `forward` / `forward_from_inputs` (:200-201, :670-677) and the prediction modes its test loop runs -
`predict_tta`, `predict_mc_dropout`, `predict_tta_mc`, `predict_custom` (:484-632, :682-702; the default test mode
is "tta_mc": 4 flips x 10 dropout passes).  Same method names, arguments and return structure; no Lightning
dependency.  Training (`_shared_step`, optimisers, losses) is not built and raises.

MC dropout follows the reference's switch exactly: `enable_dropout` puts the nn.Dropout sub-modules of the two
encoders in train mode, `set_batchnorm_eval` keeps BatchNorm frozen, and the encoders' forward then arms the
Philox dropout of the GEMM / stem epilogues.  The random stream is the kernel's own (seeded with
`set_mc_seed`), so parity with the reference is statistical, not bitwise."""
from __future__ import annotations

import torch
import torch.nn as nn

from train import tta_flip_lr, tta_flip_lrud, tta_flip_ud, tta_id

__all__ = ["LightningFusionModel"]


def _collapse_gating(gw):
    if gw.dim() == 5:
        return gw.mean(dim=(2, 3, 4))
    if gw.dim() == 4:
        return gw.mean(dim=(2, 3))
    if gw.dim() == 2:
        return gw
    raise ValueError(f"Unexpected gating weight shape: {gw.shape}")


class LightningFusionModel(nn.Module):
    def __init__(self, dwi_model, dce_model, fusion_model, parameters_dict=None):
        super().__init__()
        self.dwi_model, self.dce_model, self.fusion_model = dwi_model, dce_model, fusion_model
        self.parameters_dict = parameters_dict or {}
        self.mask_enabled = bool(getattr(dwi_model, "mask_enabled", False))
        self.transforms_list = [tta_id, tta_flip_lr, tta_flip_ud, tta_flip_lrud]

    @property
    def device(self):
        return next(self.fusion_model.parameters()).device

    # ---------------------------------------------------------------- forward ----
    def forward(self, dwi_feats, dce_feats, dwi_mask=None, dce_mask=None):
        return self.fusion_model(dwi_feats, dce_feats, dwi_mask, dce_mask)

    def forward_from_inputs(self, dwi_inputs, dce_inputs, masks=None):
        _, dwi_aux, dwi_mask_pred = self.dwi_model(dwi_inputs)
        _, dce_aux, dce_mask_pred = self.dce_model(dce_inputs)
        return self.forward(dwi_aux["raw_feats"], dce_aux["raw_feats"], dwi_mask_pred, dce_mask_pred)

    # ------------------------------------------------------------- MC dropout ----
    def enable_dropout(self, model):
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.train()

    def set_batchnorm_eval(self, model):
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.SyncBatchNorm)):
                m.eval()

    def _get_module_train_states(self, model):
        return {m: m.training for m in model.modules()}

    def _restore_module_train_states(self, model, states):
        for m, was_training in states.items():
            m.training = was_training  # per module, not recursive (nn.Module.train would undo the children)

    def mc_enable(self, model):
        self.enable_dropout(model)
        self.set_batchnorm_eval(model)

    def set_mc_seed(self, seed):
        self.dwi_model.set_mc_seed(seed)
        self.dce_model.set_mc_seed(seed + 1)

    @torch.no_grad()
    def predict_mc_dropout(self, dwi_inputs, dce_inputs, masks=None, passes=20):
        states_dwi = self._get_module_train_states(self.dwi_model)
        states_dce = self._get_module_train_states(self.dce_model)
        self.mc_enable(self.dwi_model)
        self.mc_enable(self.dce_model)
        preds, gating = [], []
        dwi_aux = dce_aux = None
        try:
            for _ in range(passes):
                _, dwi_aux, dwi_mask = self.dwi_model(dwi_inputs)
                _, dce_aux, dce_mask = self.dce_model(dce_inputs)
                logits, _, aux = self.forward(dwi_aux["raw_feats"], dce_aux["raw_feats"], dwi_mask, dce_mask)
                gw = aux["gating_weights"]
                if gw is not None:
                    gating.append(_collapse_gating(gw))
                preds.append(torch.softmax(logits, dim=1))
        finally:
            self._restore_module_train_states(self.dwi_model, states_dwi)
            self._restore_module_train_states(self.dce_model, states_dce)
        stack = torch.stack(preds, dim=0)
        mean_gating = torch.stack(gating, dim=0).mean(0).cpu() if gating else None
        return stack.mean(0), stack.std(0), {"gating_weights": mean_gating, "dwi_aux": dwi_aux, "dce_aux": dce_aux}

    # -------------------------------------------------------------------- TTA ----
    @torch.no_grad()
    def predict_tta(self, dwi_inputs, dce_inputs, masks=None, transforms=None):
        transforms = self.transforms_list if transforms is None else transforms
        preds, gating = [], []
        for t in transforms:
            logits, _, aux = self.forward_from_inputs(t(x=dwi_inputs), t(x=dce_inputs), masks)
            preds.append(torch.softmax(logits, dim=1))
            gating.append(_collapse_gating(aux["gating_weights"]))
        stack = torch.stack(preds, dim=0)
        mean_gating = torch.stack(gating, dim=0).mean(0).cpu() if gating else None
        # (the reference reads aux.get("dwi_aux") from the fusion aux, which never holds it: None)
        return stack.mean(0), stack.std(0), {"gating_weights": mean_gating, "dwi_aux": None, "dce_aux": None}

    @torch.no_grad()
    def predict_tta_mc(self, dwi_inputs, dce_inputs, masks=None, transforms=None, passes=10):
        transforms = self.transforms_list if transforms is None else transforms
        all_preds, all_gating = [], []
        last_dwi_aux = last_dce_aux = None
        for t in transforms:
            mean_preds, _, aux = self.predict_mc_dropout(t(x=dwi_inputs), t(x=dce_inputs), masks=masks, passes=passes)
            all_preds.append(mean_preds)
            all_gating.append(_collapse_gating(aux["gating_weights"]))
            last_dwi_aux, last_dce_aux = aux.get("dwi_aux"), aux.get("dce_aux")
        stack = torch.stack(all_preds, dim=0)
        mean_gating = torch.stack(all_gating, dim=0).mean(0) if all_gating else None
        return stack.mean(0), stack.std(0), {"gating_weights": mean_gating, "dwi_aux": last_dwi_aux,
                                             "dce_aux": last_dce_aux}

    def predict_custom(self, batch, mode="normal", mc_passes=10):
        dwi_inputs, dce_inputs = batch[0].to(self.device), batch[1].to(self.device)
        masks = batch[2] if len(batch) == 4 else None
        if mode == "normal":
            return self.forward_from_inputs(dwi_inputs, dce_inputs, masks)
        if mode == "tta":
            return self.predict_tta(dwi_inputs, dce_inputs, masks)
        if mode == "mc":
            return self.predict_mc_dropout(dwi_inputs, dce_inputs, passes=mc_passes)
        if mode == "tta_mc":
            return self.predict_tta_mc(dwi_inputs, dce_inputs, masks, passes=mc_passes)
        raise ValueError(f"Unknown predict mode: {mode}")

    # --------------------------------------------------------------- training ----
    def _shared_step(self, batch, phase="train", return_preds=False):
        raise NotImplementedError("the fusion training step (losses, backward, optimiser) is not built in the B200 path")

    training_step = validation_step = test_step = configure_optimizers = _shared_step
