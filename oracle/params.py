"""Seeded weights and synthetic inputs shared by the golden generator and the tests
(test infrastructure only - see oracle/__init__.py).

Weights are a pure function of (parameter name, shape, seed), so the unmodified reference
module, the oracle restatement and the CUDA product can be given bit-identical parameters
without shipping them: ``seeded_state_dict(shapes, seed)``.  BatchNorm running statistics
are randomised so that BN folding is exercised.
"""
from __future__ import annotations

import math
import zlib

import torch


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def seeded_tensor(name, shape, seed=0, dtype=torch.float32):
    g = _gen(name, seed)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == "running_var":
        return 0.5 + torch.rand(shape, generator=g)
    if len(shape) == 0:
        return 0.1 + 0.5 * torch.rand(shape, generator=g)
    if leaf.startswith("gamma"):  # LayerScale vectors, transformer_model.py:75-76
        return 0.1 * (1.0 + 0.1 * torch.randn(shape, generator=g))
    if len(shape) == 1:
        if leaf in ("bias", "in_proj_bias"):
            return 0.05 * torch.randn(shape, generator=g)
        return 1.0 + 0.1 * torch.randn(shape, generator=g)  # norm scales
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    return torch.randn(shape, generator=g) * (1.4 / math.sqrt(fan_in))


def seeded_state_dict(shapes, seed=0):
    """shapes: {name: shape}.  Returns {name: tensor}."""
    return {k: seeded_tensor(k, v, seed) for k, v in shapes.items()}


def shapes_of(state_dict):
    return {k: tuple(v.shape) for k, v in state_dict.items()}


# ------------------------------------------------------------------ inputs ----
def synthetic_raw(n, seed=1234, dwi_channels=16, dce_channels=6, size=64, kind="U"):
    """Raw ROIs before normalisation.

    kind "U": iid uniform (throughput set; no ties).  kind "S": structured - a smooth blob
    per case times a per-channel decay (DWI) / wash-in-out (DCE) curve plus 2 % noise.
    Returns (dwi_raw [n,Cd,S,S] fp32, dce_raw [n,Cc,S,S] fp32 already divided by the case
    max as prepare_single_model.py:338-339 does, mask [n,1,32,32], label [n]).
    """
    g = torch.Generator().manual_seed(seed)
    if kind == "U":
        dwi = torch.rand(n, dwi_channels, size, size, generator=g) * 1000.0 + 1.0
        dce = torch.rand(n, dce_channels, size, size, generator=g)
    else:
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing="ij")
        cx = torch.rand(n, 1, 1, generator=g) - 0.5
        cy = torch.rand(n, 1, 1, generator=g) - 0.5
        sg = 0.15 + 0.35 * torch.rand(n, 1, 1, generator=g)
        amp = 0.5 + torch.rand(n, 1, 1, generator=g)
        blob = amp * torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg ** 2)) + 0.1
        adc = 0.5 + 2.5 * torch.rand(n, 1, 1, 1, generator=g)
        b = torch.linspace(0, 1.5, dwi_channels).view(1, -1, 1, 1)
        dwi = 800.0 * blob.unsqueeze(1) * torch.exp(-b * adc) + 20.0
        dwi = dwi * (1 + 0.02 * torch.randn(n, dwi_channels, size, size, generator=g))
        t = torch.linspace(0, 1, dce_channels).view(1, -1, 1, 1)
        kin = 2.0 + 6.0 * torch.rand(n, 1, 1, 1, generator=g)
        kout = 2.0 * torch.rand(n, 1, 1, 1, generator=g)
        curve = (1 - torch.exp(-kin * t)) * torch.exp(-kout * t) + 0.05
        dce = blob.unsqueeze(1) * curve
        dce = (dce * (1 + 0.02 * torch.randn(n, dce_channels, size, size, generator=g))).clamp(min=0)
    dce = dce / dce.amax(dim=(1, 2, 3), keepdim=True)
    mask = (torch.rand(n, 1, 32, 32, generator=g) > 0.5).float()
    label = torch.randint(0, 4, (n,), generator=g)
    return dwi.float(), dce.float(), mask, label


def edge_cases(size=64, dwi_channels=16):
    """Set E for the DWI normaliser: constant plane (std clamp), all-zero case, ties, negatives, outlier."""
    g = torch.Generator().manual_seed(99)
    x = torch.rand(5, dwi_channels, size, size, generator=g) * 100.0
    x[0, 0] = 7.0                                  # constant plane -> std clamps to 1e-6
    x[1] = 0.0                                     # all-zero case
    x[2] = torch.round(x[2] / 10.0) * 10.0         # heavy ties
    x[3] = x[3] - 50.0                             # negatives
    x[4, 2, 5, 5] = 1e6                            # one huge outlier
    return x


def synthetic_head_batch(n, seed=11, channels=512, size=32, num_classes=4):
    """Inputs of the fusion-head training step: encoder f3 maps (values exactly representable in bf16, the dtype the
    product's encoders emit), encoder mask logits and labels.  Returns (f3_dwi, f3_dce [n,C,S,S], mask_dwi, mask_dce
    [n,1,S,S], labels [n] int64)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, 1, size, size, generator=g)
    f3d = (0.6 * torch.randn(n, channels, size, size, generator=g) + 0.3 * base).bfloat16().float()
    f3c = (0.6 * torch.randn(n, channels, size, size, generator=g) - 0.2 * base).bfloat16().float()
    md = torch.randn(n, 1, size, size, generator=g)
    mc = torch.randn(n, 1, size, size, generator=g)
    labels = torch.randint(0, num_classes, (n,), generator=g)
    return f3d, f3c, md, mc, labels


# ------------------------------------------------ briefly trained weights (tests/golden/trained_cnn.npz) ----
def structured_targets(dwi_raw):
    """Labels and 32 x 32 target masks that are functions of the structured ("S") ROI itself, so a few optimisation
    steps make the predictions vary over cases: class = 2 * (lesion area above the median) + (lesion centre in the
    right half); mask = the lesion (pixels above half of the first b-value image's maximum), 2 x 2 pooled."""
    d0 = dwi_raw[:, 0]
    hot = (d0 > 0.5 * d0.amax(dim=(1, 2), keepdim=True)).float()
    area = hot.mean(dim=(1, 2))
    xx = torch.linspace(-1, 1, d0.shape[-1]).view(1, 1, -1)
    w = d0 - d0.amin(dim=(1, 2), keepdim=True)
    cx = (w * xx).sum(dim=(1, 2)) / w.sum(dim=(1, 2))
    labels = 2 * (area > 0.13).long() + (cx > -0.03).long()
    masks = (torch.nn.functional.adaptive_avg_pool2d(hot.unsqueeze(1), 32) > 0.5).float()
    return labels, masks


def apply_delta(base, q, scale):
    """fixture weight = seeded + q * scale, elementwise in fp32 (bit-reproducible on any machine)."""
    return (base.float() + q.float() * torch.tensor(float(scale), dtype=torch.float32)).to(base.dtype)


def trained_state_dicts(npz, shapes, seed=7):
    """{"dwi" | "dce" | "fusion": state_dict} of the briefly trained fixture: `npz` = the loaded
    tests/golden/trained_cnn.npz (int8 deltas + per-tensor scales), `shapes` = {module: {name: shape}}."""
    out = {}
    for name, sh in shapes.items():
        sd = seeded_state_dict(sh, seed)
        for k in sd:
            key = f"{name}/{k}/q"
            if key in npz.files:
                sd[k] = apply_delta(sd[k], torch.from_numpy(npz[key]), npz[f"{name}/{k}/scale"])
            elif f"{name}/{k}/exact" in npz.files:
                sd[k] = torch.from_numpy(npz[f"{name}/{k}/exact"]).to(sd[k].dtype).reshape(sd[k].shape)
        out[name] = sd
    return out
