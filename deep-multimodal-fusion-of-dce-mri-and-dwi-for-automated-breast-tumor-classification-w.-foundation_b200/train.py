"""Test-time-augmentation transforms of the reference's train.py (lines 916-923), the part of that module that sits
on the inference path (train_fusion.LightningFusionModel.predict_tta / predict_tta_mc).  CUDA tensors are flipped by
the native kernel (b200_flip_planes); there is no CPU path.  The training loop itself is not built."""
from __future__ import annotations

import b200_native as nat

__all__ = ["tta_id", "inv_tta_id", "tta_flip_lr", "inv_tta_flip_lr", "tta_flip_ud", "inv_tta_flip_ud", "tta_flip_lrud",
           "inv_tta_flip_lrud"]


def tta_id(x):
    return x


def tta_flip_lr(x):
    return nat.flip_planes(x, True, False)


def tta_flip_ud(x):
    return nat.flip_planes(x, False, True)


def tta_flip_lrud(x):
    return nat.flip_planes(x, True, True)


inv_tta_id, inv_tta_flip_lr, inv_tta_flip_ud, inv_tta_flip_lrud = tta_id, tta_flip_lr, tta_flip_ud, tta_flip_lrud
