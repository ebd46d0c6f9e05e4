"""Inference-time surface of the reference's `LightningFusionModel` (code/train_fusion.py): the three models behind
one object, `forward` / `forward_from_inputs` (:200-201, :670-677) and the prediction modes its test loop runs -
`predict_tta`, `predict_mc_dropout`, `predict_tta_mc`, `predict_custom` (:484-632, :682-702; the default test mode
is "tta_mc": 4 flips x 10 dropout passes).  Same method names, arguments and return structure; no Lightning
dependency.  Training: `_shared_step("train")` / `training_step` / `configure_optimizers` run the fusion-head
fine-tuning step of the frozen-encoder phase - classification + mask dice terms (fusion_train.FusionHeadTrainer);
the reconstruction / mimic loss terms and unfrozen encoders are not built and raise.

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
        modes = self._aux_modes()
        try:
            for i in range(passes):
                # only the LAST pass's encoder aux is returned (train_fusion.py:536): the other passes skip every
                # output the logits do not depend on (reconstruction heads, projectors, encoder class heads)
                self._set_aux_modes(modes if i == passes - 1 else ("logits", "logits", "logits"), fusion="logits")
                _, dwi_aux, dwi_mask = self.dwi_model(dwi_inputs)
                _, dce_aux, dce_mask = self.dce_model(dce_inputs)
                logits, _, aux = self.forward(dwi_aux["raw_feats"], dce_aux["raw_feats"], dwi_mask, dce_mask)
                gw = aux["gating_weights"]
                if gw is not None:
                    gating.append(_collapse_gating(gw))
                preds.append(torch.softmax(logits, dim=1))
        finally:
            self._set_aux_modes(modes)
            self._restore_module_train_states(self.dwi_model, states_dwi)
            self._restore_module_train_states(self.dce_model, states_dce)
        stack = torch.stack(preds, dim=0)
        mean_gating = torch.stack(gating, dim=0).mean(0).cpu() if gating else None
        return stack.mean(0), stack.std(0), {"gating_weights": mean_gating, "dwi_aux": dwi_aux, "dce_aux": dce_aux}

    def _aux_modes(self):
        return tuple(getattr(m, "aux_mode", "full") for m in (self.dwi_model, self.dce_model, self.fusion_model))

    def _set_aux_modes(self, modes, fusion=None):
        for m, mode in zip((self.dwi_model, self.dce_model, self.fusion_model), modes):
            m.aux_mode = mode
        if fusion is not None:
            self.fusion_model.aux_mode = fusion

    # -------------------------------------------------------------------- TTA ----
    @torch.no_grad()
    def predict_tta(self, dwi_inputs, dce_inputs, masks=None, transforms=None):
        transforms = self.transforms_list if transforms is None else transforms
        preds, gating = [], []
        modes = self._aux_modes()
        self._set_aux_modes(("logits", "logits", "logits"))  # nothing but logits and gating weights is returned
        try:
            for t in transforms:
                logits, _, aux = self.forward_from_inputs(t(x=dwi_inputs), t(x=dce_inputs), masks)
                preds.append(torch.softmax(logits, dim=1))
                gating.append(_collapse_gating(aux["gating_weights"]))
        finally:
            self._set_aux_modes(modes)
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
    _UNBUILT_TERMS = (("recon_enabled", "reconstruction (train_fusion.py:271-284)"),
                      ("mimic_enabled", "mimic (train_fusion.py:287-292)"),
                      ("attn_reg_enabled", "attention regularisation (train_fusion.py:259-262)"))

    def _lambda_mask(self):
        mp = self.parameters_dict.get("fusion_model_parameters", {}).get("mask_parameters", {})
        if self.parameters_dict.get("b200_classification_objective_only", False) or not mp.get("mask", False):
            return 0.0
        return float(mp.get("lambda_mask", 0.0))

    def _objective_check(self):
        """The B200 step computes the classification term (train_fusion.py:238-242) and the mask dice term
        (:245-255) of the reference's total loss.  Anything else the configuration enables raises unless the caller
        opted into the classification-only objective with parameters_dict["b200_classification_objective_only"]."""
        if self.parameters_dict.get("b200_classification_objective_only", False):
            return
        fp = self.parameters_dict.get("fusion_model_parameters", {})
        on = [what for key, what in self._UNBUILT_TERMS if fp.get(key, False)]
        mp = fp.get("mask_parameters", {})
        if mp.get("mask", False) and mp.get("mask_loss_type", "dice") not in ("dice", "dice_bce"):
            raise ValueError(f"Invalid mask loss: {mp.get('mask_loss_type')}")  # selector_helpers.py:109
        if on:
            raise NotImplementedError(
                "loss terms not built in the B200 training step: " + "; ".join(on) + " - disable them or set "
                "parameters_dict['b200_classification_objective_only'] = True")

    def configure_optimizers(self):
        """The fusion-head parameter group of LightningFusionOptimizerFactory (selector_helpers.py:356-520, the only
        group in the optimiser while `backbone_freeze_on_start`), as a FusionHeadTrainer (flat buffers + fused
        AdamW).  Loss settings: label_smoothing_alpha, classification_loss_parameters.gamma; `class_weights` (the
        'wfl' inverse-frequency weights of selector_helpers.py:25-41) via set_class_weights."""
        from fusion_train import FusionHeadTrainer

        fp = self.parameters_dict.get("fusion_model_parameters", {})
        op = fp.get("optimizer_parameters", {})
        if op.get("name", "adamw").lower() != "adamw" or op.get("amsgrad", False):
            raise NotImplementedError("only AdamW without amsgrad is built")
        if not fp.get("label_smoothing_enabled", True):
            # the reference's train branch reads `smoothed` unconditionally (train_fusion.py:240-241): NameError
            raise RuntimeError("the reference training step requires label_smoothing_enabled")
        cl = fp.get("classification_loss_parameters", {})
        gamma = cl.get("gamma", None)
        # the fusion head is the LAST group of the discriminative schedule (selector_helpers.py:490-512): its learning
        # rate is the base rate (decay exponent 0) and its weight decay reg_base when discriminative_reg is on
        wd = op.get("weight_decay", 4e-5)
        if op.get("discriminative_lr", False) and op.get("discriminative_reg", False):
            wd = op.get("reg_base", wd)
        self.head_trainer = FusionHeadTrainer(
            self.fusion_model, lr=op.get("lr", 1e-4), betas=op.get("betas", (0.9, 0.999)), eps=op.get("eps", 1e-8),
            weight_decay=wd, smoothing=fp.get("label_smoothing_alpha", 0.1),
            gamma=2 if gamma is None else gamma, class_weights=getattr(self, "_class_weights", None),
            lambda_mask=self._lambda_mask(),
            mask_loss_type=fp.get("mask_parameters", {}).get("mask_loss_type", "dice"))
        sched = self._build_scheduler(fp.get("scheduler", None), self.head_trainer)
        if sched is None:
            return self.head_trainer
        return {"optimizer": self.head_trainer, "lr_scheduler": sched}   # train_fusion.py:151-161

    @staticmethod
    def _build_scheduler(cfg, optimizer):
        """LightningFusionOptimizerFactory._build_scheduler (selector_helpers.py:692-728): torch's own schedulers on
        the trainer, which is a torch.optim.Optimizer."""
        if cfg is None:
            return None
        import math

        name = cfg["name"].lower()
        if name == "reduce_lr_on_plateau":
            sch = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=cfg["factor"],
                                                             patience=cfg["patience"], min_lr=cfg["min_lr"],
                                                             threshold=cfg["threshold"])
            return {"scheduler": sch, "monitor": cfg["monitor"], "interval": "epoch"}
        if name == "cosine":
            sch = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=cfg["T_max"], eta_min=cfg["eta_min"])
            return {"scheduler": sch, "interval": "epoch"}
        if name == "cosine_with_warmup":
            warmup, max_steps = cfg.get("warmup_steps", 500), cfg.get("max_steps", 10000)

            def lr_lambda(step):
                if step < warmup:
                    return float(step) / float(warmup)
                return 0.5 * (1 + math.cos(math.pi * (step - warmup) / float(max_steps - warmup)))

            return {"scheduler": torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda), "interval": "step"}
        raise ValueError(f"Unknown scheduler: {cfg['name']}")

    def set_class_weights(self, train_labels):
        """Inverse class frequency weights of the 'wfl' loss (selector_helpers.py:25-41)."""
        counts = torch.bincount(train_labels.long().cpu(), minlength=self.fusion_model.num_classes).float()
        self._class_weights = train_labels.numel() / (len(counts) * (counts + 1e-6))
        return self._class_weights

    def _unpack(self, batch):
        masks = None
        if len(batch) == 4:
            dwi, dce, masks, labels = batch
        else:
            dwi, dce, labels = batch
        dev = self.device
        return dwi.to(dev), dce.to(dev), (masks.to(dev) if masks is not None else None), labels.long().to(dev)

    def _shared_step(self, batch, phase="train", return_preds=False):
        """train: frozen encoders (eval-mode BatchNorm, no dropout) -> fusion head forward + backward; the gradients
        are left in the trainer's flat buffer (there is no autograd graph: `optimizer_step` / `fit_batch` apply
        them).  val / test: inference forward and the hard-label loss (train_fusion.py:241)."""
        self._objective_check()
        dwi, dce, masks, labels = self._unpack(batch)
        if phase == "train":
            if getattr(self, "head_trainer", None) is None:
                self.configure_optimizers()
            if any(p.requires_grad for m in (self.dwi_model, self.dce_model) for p in m.parameters()):
                raise NotImplementedError("training with unfrozen encoders is not built (freeze them: "
                                          "backbone_freeze_on_start)")
            modes = self._aux_modes()
            self._set_aux_modes(("logits", "logits", modes[2]))  # the step needs f3 and the mask logits only
            try:
                with torch.no_grad():
                    _, dwi_aux, dwi_mask = self.dwi_model(dwi)
                    _, dce_aux, dce_mask = self.dce_model(dce)
            finally:
                self._set_aux_modes(modes)
            self.head_trainer.zero_grad()
            loss, logits = self.head_trainer.loss_and_grads(dwi_aux["raw_feats"][-1], dce_aux["raw_feats"][-1],
                                                            dwi_mask, dce_mask, labels, masks)
            loss = loss.clone().squeeze(0)
            if return_preds:
                fused_mask = getattr(self.head_trainer, "fused_mask_logits", None)
                return loss, logits.clone(), None, fused_mask.clone() if fused_mask is not None else None
            return loss
        with torch.no_grad():
            logits, fused_mask, aux = self.forward_from_inputs(dwi, dce)
            tr = getattr(self, "head_trainer", None)
            gamma = tr.gamma if tr is not None else 2.0
            lp = torch.log_softmax(logits, dim=1)  # [B, K] metric arithmetic on the logits
            fw = (1 - lp.exp()) ** gamma
            if getattr(self, "_class_weights", None) is not None:
                fw = fw * self._class_weights.to(lp.device).view(1, -1)
            onehot = torch.nn.functional.one_hot(labels, logits.shape[1]).float()
            loss = (-(onehot * fw * lp).sum(dim=1)).mean()
        if return_preds:
            return loss, logits, aux, fused_mask
        return loss

    def training_step(self, batch, batch_idx=0):
        return self._shared_step(batch, "train")

    def validation_step(self, batch, batch_idx=0):
        return self._shared_step(batch, "val")

    def test_step(self, batch, batch_idx=0):
        return self._shared_step(batch, "test")

    def optimizer_step(self):
        """Gradient all-reduce over the data-parallel ranks + AdamW; returns the rank-averaged loss."""
        return self.head_trainer.step()

    def fit_batch(self, batch):
        """One whole optimisation step: training_step + optimizer_step."""
        self.training_step(batch)
        return self.optimizer_step()
