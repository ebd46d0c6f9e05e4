"""Inference-time surface of the reference's `LightningFusionModel` (code/train_fusion.py): the three models behind
one object, `forward` / `forward_from_inputs` (:200-201, :670-677) and the prediction modes its test loop runs -
`predict_tta`, `predict_mc_dropout`, `predict_tta_mc`, `predict_custom` (:484-632, :682-702; the default test mode
is "tta_mc": 4 flips x 10 dropout passes).  Same method names, arguments and return structure; no Lightning
dependency.  Training: `_shared_step("train")` / `training_step` / `configure_optimizers` / `on_train_epoch_start` run
the reference's optimisation step - classification + three mask + three reconstruction + mimic terms
(train_fusion.py:238-296), frozen encoders (train-mode BatchNorm / dropout like Lightning's fit loop, or eval mode with
parameters_dict["b200_frozen_encoder_mode"] = "eval") or unfrozen ones with the reference's depth-wise parameter
groups - on train_graph.FullFusionTrainer (explicit backward on the training kernels); the classification + mask
objective on eval-mode frozen encoders can also run on the fp32 token-shortcut trainer fusion_train.FusionHeadTrainer.

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
    def _lambda_mask(self):
        mp = self.parameters_dict.get("fusion_model_parameters", {}).get("mask_parameters", {})
        if self.parameters_dict.get("b200_classification_objective_only", False) or not mp.get("mask", False):
            return 0.0
        return float(mp.get("lambda_mask", 0.0))

    def _objective(self):
        """Which of the reference's loss terms (train_fusion.py:238-296) the configuration turns on.  Everything the
        reference's default configuration enables is built: classification, the three mask terms, the three
        reconstruction terms and the mimic term.  Attention-energy / feature-consistency regularisation
        (`attn_reg_enabled`, off by default; it calls a 2-argument function with 3 arguments in the reference, SURVEY
        App. A-11) raises; the fusion step's feature-norm term is identically 0 in the reference (its `aux` holds no
        `raw_feats`, train.py:1021-1030)."""
        fp = self.parameters_dict.get("fusion_model_parameters", {})
        if fp.get("attn_reg_enabled", False):
            raise NotImplementedError("attn_reg_enabled: attention-energy / feature-consistency regularisation is not built")
        mp = fp.get("mask_parameters", {})
        if mp.get("mask", False) and mp.get("mask_loss_type", "dice") not in ("dice", "dice_bce"):
            raise ValueError(f"Invalid mask loss: {mp.get('mask_loss_type')}")  # selector_helpers.py:109
        cls_only = self.parameters_dict.get("b200_classification_objective_only", False)
        return {"recon": bool(fp.get("recon_enabled", False)) and not cls_only,
                "mimic": bool(fp.get("mimic_enabled", False)) and not cls_only,
                "lambda_recon": float(fp.get("lambda_recon", 0.0)), "lambda_mimic": float(fp.get("lambda_mimic", 0.0))}

    def _encoders_trainable(self):
        return any(p.requires_grad for m in (self.dwi_model, self.dce_model) for p in m.parameters())

    def _needs_full_trainer(self):
        obj = self._objective()
        mode = self.parameters_dict.get("b200_frozen_encoder_mode", "train")
        return obj["recon"] or obj["mimic"] or self._encoders_trainable() or mode == "train"

    def _group_hparams(self):
        """(lr, weight_decay) per parameter name, as LightningFusionOptimizerFactory._build_optimizer assigns them
        (selector_helpers.py:456-518): without `discriminative_lr` one group; with it the depth groups [block1],
        [block2], [block3 + everything else of the encoders], [fusion head] get lr = base / decay^(depth from the
        head) and, with `discriminative_reg`, weight decay reg_base * reg_decay^(depth from the head)."""
        fp = self.parameters_dict.get("fusion_model_parameters", {})
        op = fp.get("optimizer_parameters", {})
        base_lr, wd = op.get("lr", 1e-4), op.get("weight_decay", 4e-5)
        if not op.get("discriminative_lr", False):
            return lambda name: (base_lr, wd)
        decay, use_reg = op.get("lr_decay_factor", 2.0), op.get("discriminative_reg", False)
        reg_base, reg_decay = op.get("reg_base", wd), op.get("reg_decay_factor", 2.0)

        def fn(name):
            if name.startswith("fusion."):
                depth = 0
            elif ".block1." in name:
                depth = 3
            elif ".block2." in name:
                depth = 2
            else:
                depth = 1
            return base_lr / decay ** depth, (reg_base * reg_decay ** depth) if use_reg else wd

        return fn

    def configure_optimizers(self):
        """The optimiser of the fit: the fusion-head group alone while the encoders are frozen
        (`backbone_freeze_on_start`), every reachable parameter once they are unfrozen.  Loss settings:
        label_smoothing_alpha, classification_loss_parameters.gamma; `class_weights` (the 'wfl' inverse-frequency
        weights of selector_helpers.py:25-41) via set_class_weights.  Two implementations behind one interface:
        train_graph.FullFusionTrainer (any objective, trainable or train-mode encoders - the reference's behaviour) and
        fusion_train.FusionHeadTrainer (classification + mask terms on eval-mode frozen encoders, fp32 token shortcut;
        chosen with parameters_dict["b200_frozen_encoder_mode"] = "eval" when recon / mimic are off)."""
        fp = self.parameters_dict.get("fusion_model_parameters", {})
        op = fp.get("optimizer_parameters", {})
        if op.get("name", "adamw").lower() != "adamw" or op.get("amsgrad", False):
            raise NotImplementedError("only AdamW without amsgrad is built")
        if not fp.get("label_smoothing_enabled", True):
            # the reference's train branch reads `smoothed` unconditionally (train_fusion.py:240-241): NameError
            raise RuntimeError("the reference training step requires label_smoothing_enabled")
        cl = fp.get("classification_loss_parameters", {})
        gamma = cl.get("gamma", None)
        wd = op.get("weight_decay", 4e-5)
        if op.get("discriminative_lr", False) and op.get("discriminative_reg", False):
            wd = op.get("reg_base", wd)
        common = dict(lr=op.get("lr", 1e-4), betas=op.get("betas", (0.9, 0.999)), eps=op.get("eps", 1e-8), weight_decay=wd,
                      smoothing=fp.get("label_smoothing_alpha", 0.1), gamma=2 if gamma is None else gamma,
                      class_weights=getattr(self, "_class_weights", None), lambda_mask=self._lambda_mask())
        self.head_trainer = self.full_trainer = None
        if self._needs_full_trainer():
            from train_graph import FullFusionTrainer

            obj = self._objective()
            if fp.get("mask_parameters", {}).get("mask_loss_type", "dice") != "dice":
                raise NotImplementedError("the full objective is built with SoftDiceLoss mask terms (mask_loss_type='dice')")
            self.full_trainer = FullFusionTrainer(
                self.dwi_model, self.dce_model, self.fusion_model, lambda_recon=obj["lambda_recon"] if obj["recon"] else 0.0,
                lambda_mimic=obj["lambda_mimic"] if obj["mimic"] else 0.0, group_fn=self._group_hparams(),
                encoder_mode=self.parameters_dict.get("b200_frozen_encoder_mode", "train"), **common)
            return self.full_trainer
        from fusion_train import FusionHeadTrainer

        self.head_trainer = FusionHeadTrainer(
            self.fusion_model, mask_loss_type=fp.get("mask_parameters", {}).get("mask_loss_type", "dice"), **common)
        sched = self._build_scheduler(fp.get("scheduler", None), self.head_trainer)
        if sched is None:
            return self.head_trainer
        return {"optimizer": self.head_trainer, "lr_scheduler": sched}   # train_fusion.py:151-161

    def on_train_epoch_start(self):
        """Gradual unfreezing (train_fusion.py:155-169 -> selector_helpers.py:523-620): after the caller (or the
        reference's factory) has flipped requires_grad flags, the flat buffers are re-bound to the new trainable set."""
        tr = getattr(self, "full_trainer", None)
        if tr is not None:
            tr.aux_w = self._aux_w()
            tr.refresh()

    def _aux_w(self):
        if self.parameters_dict.get("use_simple_aux_loss_scheduling", False):
            limit = self.parameters_dict.get("aux_loss_weight_epoch_limit", 1)
            return max(0.0, 1 - getattr(self, "current_epoch", 0) / limit)
        return 1.0

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
        """train: forward + loss + explicit backward of the configured objective; the gradients are left in the
        trainer's flat buffer (there is no autograd graph: `optimizer_step` / `fit_batch` apply them).  val / test:
        inference forward and the hard-label loss (train_fusion.py:241)."""
        self._objective()
        dwi, dce, masks, labels = self._unpack(batch)
        if phase == "train":
            if getattr(self, "head_trainer", None) is None and getattr(self, "full_trainer", None) is None:
                self.configure_optimizers()
            if self.full_trainer is not None:
                tr = self.full_trainer
                if tr.encoders_trainable != self._encoders_trainable():
                    tr.refresh()
                if masks is None:
                    raise ValueError("the fusion objective needs the target masks (mask_enabled)")
                tr.aux_w = self._aux_w()
                tr.zero_grad()
                total, _ = tr.forward_backward(dwi, dce, masks, labels)
                loss = total.clone().squeeze(0)
                if return_preds:
                    return loss, tr.logits.clone(), None, tr.fused_mask_logits.unsqueeze(1).clone()
                return loss
            if self._encoders_trainable():
                raise NotImplementedError("unfrozen encoders need the full trainer (configure_optimizers picks it)")
            modes = self._aux_modes()
            self._set_aux_modes(("logits", "logits", modes[2]))  # the step needs f3 and the mask logits only
            states = [self._get_module_train_states(m) for m in (self.dwi_model, self.dce_model)]
            try:
                self.dwi_model.eval(), self.dce_model.eval()
                with torch.no_grad():
                    _, dwi_aux, dwi_mask = self.dwi_model(dwi)
                    _, dce_aux, dce_mask = self.dce_model(dce)
            finally:
                self._set_aux_modes(modes)
                for m, st in zip((self.dwi_model, self.dce_model), states):
                    self._restore_module_train_states(m, st)
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
            ft = getattr(self, "full_trainer", None)
            gamma = tr.gamma if tr is not None else (ft.loss_hp["gamma"] if ft is not None else 2.0)
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
        tr = self.full_trainer if getattr(self, "full_trainer", None) is not None else self.head_trainer
        return tr.step()

    def fit_batch(self, batch):
        """One whole optimisation step: training_step + optimizer_step."""
        self.training_step(batch)
        return self.optimizer_step()
