"""Batched dual-modality classification pipeline: raw ROIs -> logits.

This is the B200 arrangement of the reference's inference path: per-sample CPU normalisation
in the dataset (code/dataset.py:70-98) + `fusion_model_test`'s per-batch loop
(code/model_test.py:114-155) / `LightningFusionModel.forward_from_inputs`
(code/train_fusion.py:670-677) become one device-resident step:

    DWINormalize.batch, DCENormalize.batch  ->  dwi_model, dce_model  ->  fusion_model

The per-plane means the normaliser kernels emit feed the encoders' modality SE block, so the
normalised inputs are read once.  `classify_host` is the end-to-end entry: pinned host
tensors in, host logits out, with the next batch's upload overlapped on a copy stream.
"""
from __future__ import annotations

import torch

import b200_native as nat
from dataset import DCENormalize, DWINormalize, Resize


class FusionPipeline:
    def __init__(self, dwi_model, dce_model, fusion_model, nyul_standardizer, dwi_normalize=None, aux_mode="full",
                 input_size=None, fuse_normalise=True):
        """`input_size`: the reference's `transforms.Resize(input_size)` ahead of the normalisers
        (code/prepare_single_model.py:112-120) - 224 for the ViT-B/16 encoders (C4), None = keep the ROI size.
        `fuse_normalise`: CNN encoders read the RAW ROIs and normalise inside their first layer's operand load
        (statistics / table kernels + b200_stem_ex; the normalised tensors are never written).  Bit-identical to the
        unfused path; backbone encoders (which consume the normalised tensor as a whole) keep the stand-alone passes."""
        self.dwi_model, self.dce_model, self.fusion_model = dwi_model, dce_model, fusion_model
        self.resize = Resize(input_size) if input_size is not None else None
        self.dwi_norm = dwi_normalize if dwi_normalize is not None else DWINormalize()
        self.dce_norm = DCENormalize(nyul_standardizer)
        self.set_aux_mode(aux_mode)
        self._copy_stream = None
        self.fuse_normalise = bool(fuse_normalise)

    def _can_fuse(self, dwi_raw):
        def ok(m):
            return (not m.use_backbone and m.modality_attention is not None and not m.training)

        n = dwi_raw.shape[-1] * dwi_raw.shape[-2]
        return (self.fuse_normalise and self.resize is None and ok(self.dwi_model) and ok(self.dce_model) and
                n % 4 == 0 and n <= 8192)

    def _encode(self, dwi_raw, dce_raw):
        """Normalise + both encoders -> (out_dwi, out_dce); fused first layer when the configuration allows it."""
        B = dwi_raw.shape[0]
        dev = dwi_raw.device
        if self.resize is not None:
            dwi_raw, dce_raw = self.resize.batch(dwi_raw), self.resize.batch(dce_raw)
        if self._can_fuse(dwi_raw):
            dwi_raw, dce_raw = dwi_raw.contiguous().float(), dce_raw.contiguous().float()
            norm_d, pm_d = self.dwi_norm.fused_params(dwi_raw)
            norm_c, pm_c = self.dce_norm.fused_params(dce_raw)
            return (self.dwi_model(dwi_raw, None, plane_mean=pm_d, input_norm=norm_d),
                    self.dce_model(dce_raw, None, plane_mean=pm_c, input_norm=norm_c))
        pm_d = torch.empty(B * dwi_raw.shape[1], dtype=torch.float32, device=dev)
        pm_c = torch.empty(B * dce_raw.shape[1], dtype=torch.float32, device=dev)
        dwi = self.dwi_norm.batch(dwi_raw, plane_mean=pm_d)
        dce = self.dce_norm.batch(dce_raw, plane_mean=pm_c)
        return self.dwi_model(dwi, None, plane_mean=pm_d), self.dce_model(dce, None, plane_mean=pm_c)

    def set_aux_mode(self, mode):
        if mode not in ("full", "logits"):
            raise ValueError("aux_mode must be 'full' or 'logits'")
        self.aux_mode = mode
        for m in (self.dwi_model, self.dce_model, self.fusion_model):
            m.aux_mode = mode

    def eval(self):
        for m in (self.dwi_model, self.dce_model, self.fusion_model):
            m.eval()
        return self

    @torch.no_grad()
    def forward_raw(self, dwi_raw, dce_raw, return_all=False):
        """dwi_raw [B,Cd,H,W], dce_raw [B,Cc,H,W] fp32 CUDA (DCE already divided by the case max,
        code/prepare_single_model.py:338-339).  Returns fusion logits [B,K] (fp32)."""
        out_d, out_c = self._encode(dwi_raw, dce_raw)
        out_f = self.fusion_model(out_d[1]["raw_feats"], out_c[1]["raw_feats"], out_d[2], out_c[2])
        if return_all:
            return out_d, out_c, out_f
        return out_f[0]

    def _staged(self, batches, dev):
        """Yield device copies of the host batches; batch i+1 is uploaded on a copy stream while batch i is in use."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream(dev)

        def upload(items):
            with torch.cuda.stream(self._copy_stream):
                out = tuple(t.to(dev, non_blocking=True) for t in items)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return out, ev

        it = iter(batches)
        try:
            staged = upload(next(it))
        except StopIteration:
            return
        while staged is not None:
            tensors, ev = staged
            try:
                nxt = upload(next(it))
            except StopIteration:
                nxt = None
            compute.wait_event(ev)
            for t in tensors:
                t.record_stream(compute)
            yield tensors
            staged = nxt

    @torch.no_grad()
    def classify_host(self, batches, device="cuda"):
        """End-to-end: iterable of (dwi_host, dce_host) PINNED fp32 CPU tensors -> list of host logits.
        Upload of batch i+1 runs on a copy stream while batch i computes."""
        dev = torch.device(device)
        results = []
        pool, n = None, (len(batches) if hasattr(batches, "__len__") else 0)
        for i, (d, c) in enumerate(self._staged(batches, dev)):
            logits = self.forward_raw(d, c)
            if i == 0 and n > 0:  # one pinned allocation for the whole sequence instead of one per step
                pool = torch.empty((n,) + tuple(logits.shape), dtype=logits.dtype, pin_memory=True)
            host = pool[i] if pool is not None and i < n and pool.shape[1:] == logits.shape else \
                torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
            host.copy_(logits, non_blocking=True)
            results.append(host)
        torch.cuda.current_stream(dev).synchronize()
        return results

    @torch.no_grad()
    def encode_raw(self, dwi_raw, dce_raw):
        """Normalisers + the two (frozen) encoders: -> (f3_dwi, f3_dce, dwi_mask_pred, dce_mask_pred), the inputs of
        the fusion head (code/train_fusion.py:226-236)."""
        out_d, out_c = self._encode(dwi_raw, dce_raw)
        return out_d[1]["raw_feats"][-1], out_c[1]["raw_feats"][-1], out_d[2], out_c[2]

    def fit_host(self, batches, trainer, device="cuda"):
        """End-to-end fine-tuning of the fusion head: iterable of (dwi_host, dce_host, labels_host[, masks_host])
        PINNED CPU tensors, one optimisation step per batch with `trainer` (fusion_train.FusionHeadTrainer: forward,
        backward, gradient all-reduce over the data-parallel ranks, AdamW).  Returns the per-step rank-averaged
        losses as pinned host tensors; the upload of batch i+1 overlaps the step on batch i."""
        dev = torch.device(device)
        losses = []
        for items in self._staged(batches, dev):
            d, c, lab = items[:3]
            masks = items[3] if len(items) > 3 else None
            loss, _ = trainer.train_step(*self.encode_raw(d, c), lab, masks)
            host = torch.empty(1, dtype=torch.float32, pin_memory=True)
            host.copy_(loss, non_blocking=True)
            losses.append(host)
        torch.cuda.current_stream(dev).synchronize()
        return losses
