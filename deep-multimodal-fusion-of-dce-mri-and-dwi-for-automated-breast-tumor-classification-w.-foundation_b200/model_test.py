"""B200-native drop-in for the reference's ``model_test.fusion_model_test`` (code/model_test.py:99-202).

Same signature and printed summary; the per-batch arithmetic is the three B200 modules'
eval forward (logits-only work is enough for accuracy, so aux outputs are skipped here)."""
from __future__ import annotations

import time

import torch


def fusion_model_test(dwi_model, dce_model, fusion_model, dataloaders, device, mask_fusion=False):
    since = time.time()
    for m in (fusion_model, dwi_model, dce_model):
        m.eval()
    saved = [m.aux_mode for m in (dwi_model, dce_model, fusion_model)]
    for m in (dwi_model, dce_model, fusion_model):
        m.aux_mode = "logits"
    correct = torch.zeros((), dtype=torch.int64, device=device)
    try:
        with torch.no_grad():
            for dwi_inputs, dce_inputs, masks_batch, labels in dataloaders["test"]:
                dwi_inputs = dwi_inputs.to(device, non_blocking=True)
                dce_inputs = dce_inputs.to(device, non_blocking=True)
                labels = labels.to(device, non_blocking=True)
                _, dwi_aux, dwi_mask = dwi_model(dwi_inputs, masks_batch)
                _, dce_aux, dce_mask = dce_model(dce_inputs, masks_batch)
                logits, _, _ = fusion_model(dwi_aux["raw_feats"], dce_aux["raw_feats"], dwi_mask, dce_mask)
                correct += (logits.argmax(dim=1) == labels.view(-1)).sum()  # stays on device: no per-batch sync
    finally:
        for m, mode in zip((dwi_model, dce_model, fusion_model), saved):
            m.aux_mode = mode
    n = len(dataloaders["test"].dataset)
    acc = correct.item() / n if n > 0 else 0.0
    print(f"Test Acc: {acc:.4f}")
    elapsed = time.time() - since
    print(f"Testing complete in {elapsed // 60:.0f}m {elapsed % 60:.0f}s")
    torch.cuda.empty_cache()
    return acc
