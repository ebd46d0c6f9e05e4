"""On-disk formats of the reference's trained weights, host side only.

The reference writes its trained models in two layouts, and both load into the B200 modules unchanged because these
keep the reference's parameter names and shapes:

* the "old style" dictionary `fusion_model_dict.pth` (code/run_training.py:316-326): one `torch.save`d dict with the
  entries `fusion_{fold}`, `dwi_{fold}`, `dce_{fold}`, each a plain `state_dict()`; re-saving merges into the file;
* Lightning checkpoints `best-v{n}.ckpt` (code/run_training.py:93-127, :203-270; read back at
  code/prepare_single_model.py:208-216): a dict whose `state_dict` entry holds the LightningModule's parameters -
  prefixed `model.` for `LightningSingleModel`, `dwi_model.` / `dce_model.` / `fusion_model.` for
  `LightningFusionModel` (plus metric / criterion buffers, which are skipped).

Nothing here touches the GPU; tensors are loaded to the CPU and copied by `load_state_dict`.
"""
from __future__ import annotations

import os

import torch

__all__ = ["save_model_dict", "load_model_dict", "split_lightning_state_dict", "load_lightning_checkpoint",
           "apply_states"]

_FUSION_PREFIXES = {"dwi": "dwi_model.", "dce": "dce_model.", "fusion": "fusion_model."}


def save_model_dict(path, fold, dwi_model, dce_model, fusion_model):
    """code/run_training.py:316-326: merge this fold's three state dicts into the dictionary file at `path`."""
    model_dict = torch.load(path, map_location="cpu") if os.path.exists(path) else {}
    for name, m in (("fusion", fusion_model), ("dwi", dwi_model), ("dce", dce_model)):
        model_dict[f"{name}_{fold}"] = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    torch.save(model_dict, path)
    return model_dict


def load_model_dict(path, fold):
    """-> {"dwi": state_dict, "dce": state_dict, "fusion": state_dict} of one fold."""
    model_dict = torch.load(path, map_location="cpu")
    out = {}
    for name in ("dwi", "dce", "fusion"):
        key = f"{name}_{fold}"
        if key not in model_dict:
            raise KeyError(f"{path} holds no entry '{key}' (entries: {sorted(model_dict)})")
        out[name] = model_dict[key]
    return out


def split_lightning_state_dict(state_dict):
    """Split a LightningModule state dict by model.  LightningFusionModel -> {"dwi", "dce", "fusion"};
    LightningSingleModel (prefix `model.`) -> {"model"}.  Entries of other sub-modules (torchmetrics, criteria) are
    dropped."""
    out = {}
    for name, prefix in _FUSION_PREFIXES.items():
        part = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if part:
            out[name] = part
    if not out:
        part = {k[len("model."):]: v for k, v in state_dict.items() if k.startswith("model.")}
        if part:
            out["model"] = part
    if not out:
        raise ValueError("no 'model.', 'dwi_model.', 'dce_model.' or 'fusion_model.' entries in the state dict")
    return out


def load_lightning_checkpoint(path):
    """code/prepare_single_model.py:213-214: `ckpt['state_dict'] if 'state_dict' in ckpt else ckpt`, then split."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    return split_lightning_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)


def apply_states(states, strict=True, **models):
    """apply_states(load_model_dict(p, fold), dwi=dwi_model, dce=dce_model, fusion=fusion_model): load_state_dict on
    every model named; returns the (missing, unexpected) key lists per model.  `strict=False` reproduces the
    reference's tolerant load (code/prepare_single_model.py:215)."""
    report = {}
    for name, model in models.items():
        if name not in states:
            raise KeyError(f"no state for '{name}' (have {sorted(states)})")
        res = model.load_state_dict(states[name], strict=strict)
        report[name] = (list(res.missing_keys), list(res.unexpected_keys))
    return report
