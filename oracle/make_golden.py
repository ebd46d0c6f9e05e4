"""Generates tests/golden/* by running the UNMODIFIED reference modules (authoring container only).

    python oracle/make_golden.py            # needs /root/reference (read-only)

The reference cannot travel to the GPU box, so its outputs on seeded inputs and seeded
weights are committed as small fixtures: full small tensors (logits, gates, masks) and
strided probes + sums of the large feature maps.  Inputs and weights are NOT stored - they
are regenerated from seeds by oracle/params.py.  Nothing here is imported by the product.
"""
from __future__ import annotations

import copy
import json
import os
import runpy
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/code"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import params as op  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PROBE_STRIDE = 997
PROBE_N = 256


def probe(t):
    """Compact fingerprint of a tensor: strided samples, sum, sum of |x|."""
    f = t.detach().double().reshape(-1)
    return {"shape": list(t.shape), "samples": f[::PROBE_STRIDE][:PROBE_N].float().numpy(),
            "sum": float(f.sum()), "abs_sum": float(f.abs().sum())}


def reference_parameters():
    """Execute code/parameters_generate.py with torch.save stubbed (it writes to Drive paths)."""
    saved = torch.save
    torch.save = lambda *a, **k: None
    try:
        ns = runpy.run_path(os.path.join(REF, "parameters_generate.py"))
    finally:
        torch.save = saved
    p = copy.deepcopy(ns["parameters"])
    for m in ("dwi", "dce", "fusion"):  # break the aliasing (parameters_generate.py:174, :183)
        p[f"{m}_model_parameters"] = copy.deepcopy(p[f"{m}_model_parameters"])
    return p


def configure(p, hybrid=False, input_size=64, downsample=None, mask_stage=None, repeat_blocks=None):
    p["dwi_channel_num"], p["dce_channel_num"] = 16, 6
    for m in ("dwi", "dce", "fusion"):
        mp = p[f"{m}_model_parameters"]
        mp["use_backbone"] = False
        mp["input_size"] = input_size
        mp["use_hybrid_transformer"] = hybrid and m != "fusion"
        if downsample is not None and m != "fusion":
            mp["downsample"] = downsample
        if mask_stage is not None and m != "fusion":
            mp["mask_parameters"] = dict(mp["mask_parameters"], mask_stage=mask_stage)
        if repeat_blocks is not None and m != "fusion":
            mp["repeat_blocks"] = repeat_blocks
    return p


def flatten(prefix, obj, out):
    if obj is None:
        return
    if torch.is_tensor(obj):
        pr = probe(obj)
        out[prefix + "/samples"] = pr["samples"]
        out[prefix + "/meta"] = np.array(pr["shape"] + [0], dtype=np.float64)
        out[prefix + "/sums"] = np.array([pr["sum"], pr["abs_sum"]])
        if obj.numel() <= 8192:
            out[prefix + "/full"] = obj.detach().float().numpy()
    elif isinstance(obj, (list, tuple)):
        for i, o in enumerate(obj):
            flatten(f"{prefix}.{i}", o, out)
    elif isinstance(obj, dict):
        for k, o in obj.items():
            flatten(f"{prefix}.{k}", o, out)


def model_goldens(tag, hybrid, input_size=64, downsample=None, kinds=("U", "S"), mask_stage=None, repeat_blocks=None):
    import model_module as mm  # the reference module, imported from /root/reference/code

    p = configure(reference_parameters(), hybrid, input_size, downsample, mask_stage, repeat_blocks)
    torch.manual_seed(0)
    models = {"dwi": mm.ModelMaskHeadBackbone("dwi", p, None), "dce": mm.ModelMaskHeadBackbone("dce", p, None),
              "fusion": mm.FusionModel(p)}
    shapes = {}
    for name, m in models.items():
        sh = op.shapes_of(m.state_dict())
        shapes[name] = {k: list(v) for k, v in sh.items()}
        m.load_state_dict(op.seeded_state_dict(sh, seed=7))
        m.eval()
    out = {}
    for kind in kinds:
        dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, size=input_size, kind=kind)
        # model inputs are the normalised tensors; any fp32 tensor in [0,1] does for model parity
        dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
        dce = dce_raw
        with torch.no_grad():
            ld, ad, md = models["dwi"](dwi)
            lc, ac, mc = models["dce"](dce)
            lf, mf, af = models["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
        flatten(f"{kind}/dwi/logits", ld, out)
        flatten(f"{kind}/dwi/aux", ad, out)
        flatten(f"{kind}/dwi/mask", md, out)
        flatten(f"{kind}/dce/logits", lc, out)
        flatten(f"{kind}/dce/aux", ac, out)
        flatten(f"{kind}/dce/mask", mc, out)
        flatten(f"{kind}/fusion/logits", lf, out)
        flatten(f"{kind}/fusion/mask", mf, out)
        flatten(f"{kind}/fusion/aux", af, out)
    np.savez_compressed(os.path.join(GOLD, f"model_{tag}.npz"), **out)
    if input_size == 64 and downsample is None:  # the geometry variants share the parameter shapes of "cnn"
        with open(os.path.join(GOLD, f"state_shapes_{tag}.json"), "w") as f:
            json.dump(shapes, f, indent=0, sort_keys=True)
    print(tag, "fusion logits", lf)


class _StandInInfo:
    def __init__(self, channels, reduction):
        self._c, self._r = channels, reduction

    def channels(self):
        return self._c

    def reduction(self):
        return self._r


def standin_vit(in_chans, depth=12):
    """timm is absent here (and un-pinned in the reference): the backbone handed to the reference's
    ModelMaskHeadBackbone is a parameter container with timm's ViT key names whose forward is the restated
    oracle/backbone_oracle.vit_features (itself cross-checked against torchvision's VisionTransformer)."""
    import torch.nn as nn
    from oracle import backbone_oracle as bo

    class StandInViT(nn.Module):
        def __init__(self):
            super().__init__()
            for key, shape in bo.vit_shapes(in_chans, depth=depth).items():
                mod = self
                *path, leaf = key.split(".")
                for name in path:
                    if not hasattr(mod, name):
                        mod.add_module(name, nn.Module())
                    mod = getattr(mod, name)
                mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape)))
            self.feature_info = _StandInInfo([768] * depth, [16] * depth)

        def forward(self, x):
            return bo.vit_features(dict(self.state_dict()), x)

    return StandInViT()


def configure_vit(p):
    """What foundation_model.build_medical_backbone writes for the ViT branch (foundation_model.py:526-545),
    plus the fusion input width that has to follow by hand (SURVEY.md note 9)."""
    p["dwi_channel_num"], p["dce_channel_num"] = 16, 6
    for m in ("dwi", "dce", "fusion"):
        mp = p[f"{m}_model_parameters"]
        mp["input_size"] = 224
        mp["use_hybrid_transformer"] = False
        mp["use_backbone"] = m != "fusion"
        if m != "fusion":
            mp["backbone_index_lists"] = [[0, 1, 2], [3, 4, 5, 6], [7, 8, 9, 10, 11]]
            mp["downsample"] = (False, False, False)
            mp["channels"] = (768, 768, 768)
            mp["transformer_backbone"] = True
    fs = p["fusion_model_parameters"]["fusion_specific_parameters"]
    fs["dwi_out_channels"] = fs["dce_out_channels"] = 768
    return p


def vit_goldens():
    """C4: both encoders with the ViT-B/16 backbone adapter at 224 x 224 + the fusion head on 14 x 14 maps."""
    import model_module as mm

    p = configure_vit(reference_parameters())
    torch.manual_seed(0)
    models = {"dwi": mm.ModelMaskHeadBackbone("dwi", p, standin_vit(16)),
              "dce": mm.ModelMaskHeadBackbone("dce", p, standin_vit(6)), "fusion": mm.FusionModel(p)}
    shapes = {}
    for name, m in models.items():
        sh = op.shapes_of(m.state_dict())
        shapes[name] = {k: list(v) for k, v in sh.items()}
        m.load_state_dict(op.seeded_state_dict(sh, seed=11))
        m.eval()
    out = {}
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=4321, size=224, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    dce = dce_raw
    with torch.no_grad():
        ld, ad, md = models["dwi"](dwi)
        lc, ac, mc = models["dce"](dce)
        lf, mf, af = models["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
    for tag, val in (("dwi/logits", ld), ("dwi/aux", ad), ("dwi/mask", md), ("dce/logits", lc), ("dce/aux", ac),
                     ("dce/mask", mc), ("fusion/logits", lf), ("fusion/mask", mf), ("fusion/aux", af)):
        flatten(f"S/{tag}", val, out)
    np.savez_compressed(os.path.join(GOLD, "model_vit.npz"), **out)
    with open(os.path.join(GOLD, "state_shapes_vit.json"), "w") as f:
        json.dump(shapes, f, indent=0, sort_keys=True)
    print("vit fusion logits", lf)


def standin_resnet(in_chans):
    """timm is absent: the ResNet-50 feature extractor handed to the reference is torchvision's ResNet (same
    architecture and parameter names as timm's resnet50; dilated to output stride 8 like the reference's
    `output_stride=8`), with the classifier removed and `feature_info` added."""
    import torch.nn as nn
    import torchvision

    class StandInResNet(torchvision.models.ResNet):
        def __init__(self):
            super().__init__(torchvision.models.resnet.Bottleneck, [3, 4, 6, 3],
                             replace_stride_with_dilation=[False, True, True])
            self.conv1 = nn.Conv2d(in_chans, 64, 7, stride=2, padding=3, bias=False)
            del self.fc
            self.feature_info = _StandInInfo([256, 512, 1024, 2048], [4, 8, 8, 8])

        def forward(self, x):
            t = self.maxpool(self.relu(self.bn1(self.conv1(x))))
            feats = []
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                t = layer(t)
                feats.append(t)
            return feats

    return StandInResNet()


def configure_resnet(p):
    """What foundation_model.build_medical_backbone writes for the radimagenet / resnet50 branches (:503-524, :547-569)."""
    p["dwi_channel_num"], p["dce_channel_num"] = 16, 6
    for m in ("dwi", "dce", "fusion"):
        mp = p[f"{m}_model_parameters"]
        mp["input_size"] = 224
        mp["use_hybrid_transformer"] = False
        mp["use_backbone"] = m != "fusion"
        if m != "fusion":
            mp["backbone_index_lists"] = [[0], [1], [2, 3]]
            mp["downsample"] = (True, False, False)
            mp["downsample_each_repeat"] = False
            mp["transformer_backbone"] = False
    return p


def resnet_goldens():
    """The reference's default backbone family: ResNet-50 (RadImageNet) encoders at 224 x 224 + the fusion head."""
    import model_module as mm

    p = configure_resnet(reference_parameters())
    torch.manual_seed(0)
    models = {"dwi": mm.ModelMaskHeadBackbone("dwi", p, standin_resnet(16)),
              "dce": mm.ModelMaskHeadBackbone("dce", p, standin_resnet(6)), "fusion": mm.FusionModel(p)}
    shapes = {}
    for name, m in models.items():
        sh = op.shapes_of(m.state_dict())
        shapes[name] = {k: list(v) for k, v in sh.items()}
        m.load_state_dict(op.seeded_state_dict(sh, seed=13))
        m.eval()
    out = {}
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=4321, size=224, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    with torch.no_grad():
        ld, ad, md = models["dwi"](dwi)
        lc, ac, mc = models["dce"](dce_raw)
        lf, mf, af = models["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
    for tag, val in (("dwi/logits", ld), ("dwi/aux", ad), ("dwi/mask", md), ("dce/logits", lc), ("dce/aux", ac),
                     ("dce/mask", mc), ("fusion/logits", lf), ("fusion/mask", mf), ("fusion/aux", af)):
        flatten(f"S/{tag}", val, out)
    np.savez_compressed(os.path.join(GOLD, "model_resnet.npz"), **out)
    with open(os.path.join(GOLD, "state_shapes_resnet.json"), "w") as f:
        json.dump(shapes, f, indent=0, sort_keys=True)
    print("resnet fusion logits", lf, [tuple(t.shape) for t in ad["raw_feats"]])


def normalizer_goldens():
    import dataset as ref_dataset
    import preprocess_helpers as ref_pre

    out = {}
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    dwi_u, dce_u, _, _ = op.synthetic_raw(3, seed=77, kind="U")
    norm = ref_dataset.DWINormalize()
    for name, x in (("S", dwi_raw[:3]), ("U", dwi_u), ("E", op.edge_cases())):
        y = torch.stack([norm(c) for c in x])
        flatten(f"dwi/{name}", y, out)
    y = torch.stack([ref_dataset.DWINormalize(clip_z=(-2, 2.5), adc=False)(c) for c in dwi_u])
    flatten("dwi/U_noadc", y, out)
    import contextlib
    import io
    nyul = ref_pre.NyulStandardizer()
    with contextlib.redirect_stdout(io.StringIO()):
        nyul.fit(list(dce_raw[:8]), num_channels=6)
    lm = np.stack([nyul.channel_landmarks[c] for c in range(6)])
    out["nyul/landmarks"] = lm
    for name, x in (("S", dce_raw[8:]), ("U", dce_u)):
        y = torch.stack([ref_dataset.DCENormalize(nyul)(c) for c in x])
        flatten(f"nyul/{name}", y, out)
    ties = torch.round(dce_u * 20) / 20  # tied percentiles
    y = torch.stack([nyul.transform(c) for c in ties])
    flatten("nyul/ties", y, out)
    adc = ref_pre.compute_adc_map(dwi_raw[0, :13], list(range(13)))
    out["adc/map"] = adc.numpy()
    np.savez_compressed(os.path.join(GOLD, "normalizers.npz"), **out)
    print("normalizers done")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    normalizer_goldens()
    model_goldens("cnn", hybrid=False)
    model_goldens("hybrid", hybrid=True)
    # non-default geometries: 128 x 128 ROIs (64 x 64 maps, strided mask head, 2x2-averaging projector pool) and a
    # stride-2 block3 (16 x 16 f3, fusion head with the bilinear mask path)
    model_goldens("cnn128", hybrid=False, input_size=128, kinds=("S",))
    model_goldens("cnn_s2", hybrid=False, downsample=(True, False, True), kinds=("S",))
    # mask head on f1 / f3 and repeated bottlenecks
    model_goldens("cnn_mf1", hybrid=False, kinds=("S",), mask_stage="f1")
    model_goldens("cnn_mf3", hybrid=False, kinds=("S",), mask_stage="f3")
    model_goldens("cnn_r2", hybrid=False, kinds=("S",), repeat_blocks=(2, 1, 2))
    vit_goldens()
    resnet_goldens()
