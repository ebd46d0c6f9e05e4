"""B200-native drop-in for the reference's ``prepare_fusion_model`` (code/prepare_fusion_model.py).

`prepare_fusion_model` assembles the cached per-fold tensors into loaders and instantiates the
(B200) FusionModel; `custom_double_input_collate_fn` batches (dwi, dce[, mask], label) tuples."""
from __future__ import annotations

import os

import torch
import torch.utils.data
from torch.utils.data._utils.collate import default_collate

from dataset import LoadedFusionDataset
from model_module import FusionModel


def load_dataset_split(load_path):
    return torch.load(load_path)


def custom_double_input_collate_fn(batch):
    """-> (dwi[B,..], dce[B,..], masks[B,..] | None, labels[B]) (reference :88-113)."""
    cols = {3: ([], [], None, []), 4: ([], [], [], [])}
    n = {len(item) for item in batch}
    if not n <= {3, 4}:
        raise RuntimeError("Dataset item must be 3 or 4 elements")
    dwi, dce, masks, labels = [], [], [], []
    for item in batch:
        if len(item) == 4:
            d, c, m, y = item
            masks.append(m)
        else:
            d, c, y = item
        dwi.append(d)
        dce.append(c)
        labels.append(y)
    return (default_collate(dwi), default_collate(dce), default_collate(masks) if masks else None,
            default_collate(labels))


def prepare_fusion_model(dwi_results, dce_results, fold, parameters, device, method="fusion"):
    """Loads `<data_path>/{dwi,dce}{fold}{split}data`, builds loaders (batch_size from the parameter
    dict, never shuffled - the reference compares a list with a str at :63) and a FusionModel."""
    names = parameters["namelist"]
    mask_on = parameters[f"{method}_model_parameters"]["mask_parameters"]["mask"]
    loaders = {}
    for split in names:
        dwi = load_dataset_split(os.path.join(parameters["data_path"], f"dwi{fold}{split}data"))
        dce = load_dataset_split(os.path.join(parameters["data_path"], f"dce{fold}{split}data"))
        masks = dwi["masks"] if (mask_on and dwi["masks"] is not None) else None
        ds = LoadedFusionDataset(dwi=dwi["imgs"], dce=dce["imgs"], masks=masks, labels=dwi["labels"])
        loaders[split] = torch.utils.data.DataLoader(ds, batch_size=parameters["batch_size"], shuffle=False,
                                                     num_workers=0, drop_last=False, pin_memory=True,
                                                     collate_fn=custom_double_input_collate_fn)
    return loaders, dwi_results["trained_model"], dce_results["trained_model"], FusionModel(parameters)
