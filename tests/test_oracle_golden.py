"""CPU: the oracle restatement against the fixtures produced by the unmodified reference
(oracle/make_golden.py).  Tolerances: the model restatement reorders no arithmetic, so it
must agree to fp32 round-off (1e-5 of the tensor's max); normalisers to 1e-6."""
import numpy as np
import pytest
import torch

import golden_util as gu
import parameters_default as pd
from oracle import model_oracle as mo
from oracle import normalize_oracle as no
from oracle import params as op


def _params(hybrid):
    p = pd.default_parameters()
    for m in ("dwi", "dce"):
        p[f"{m}_model_parameters"]["use_hybrid_transformer"] = hybrid
    return p


@pytest.mark.parametrize("tag,hybrid", [("cnn", False), ("hybrid", True)])
def test_model_oracle_matches_reference(tag, hybrid):
    gold = gu.load(f"model_{tag}.npz")
    shapes = gu.load_shapes(tag)
    p = _params(hybrid)
    sds = {m: op.seeded_state_dict(shapes[m], seed=7) for m in ("dwi", "dce", "fusion")}
    torch.set_num_threads(8)
    for kind in ("U", "S"):
        dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, kind=kind)
        dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
        with torch.no_grad():
            ld, ad, md = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
            lc, ac, mc = mo.encoder_forward(sds["dce"], "dce", p, dce_raw)
            lf, mf, af = mo.fusion_forward(sds["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)
        outs = {f"{kind}/dwi/logits": ld, f"{kind}/dwi/aux": ad, f"{kind}/dwi/mask": md,
                f"{kind}/dce/logits": lc, f"{kind}/dce/aux": ac, f"{kind}/dce/mask": mc,
                f"{kind}/fusion/logits": lf, f"{kind}/fusion/mask": mf, f"{kind}/fusion/aux": af}
        n = 0
        for prefix, obj in outs.items():
            for key, t in gu.walk(prefix, obj):
                gu.check(gold, key, t, rtol=2e-5)
                n += 1
        assert n == 34


def vit_parameters():
    """C4 configuration: foundation_model.build_medical_backbone rewrites the encoder parameters exactly as the
    reference does (foundation_model.py:526-545); the fusion input widths follow by hand (SURVEY.md note 9)."""
    import foundation_model as fm

    p = pd.default_parameters(input_size=224)
    backbones = {}
    for m, c in (("dwi", 16), ("dce", 6)):
        mp = p[f"{m}_model_parameters"]
        mp["backbone_str"], mp["use_backbone"] = "vit_base_patch16_224", True
        backbones[m] = fm.build_medical_backbone(p, None, m, in_channels=c)
    fs = p["fusion_model_parameters"]["fusion_specific_parameters"]
    fs["dwi_out_channels"] = fs["dce_out_channels"] = 768
    return p, backbones


def vit_inputs():
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=4321, size=224, kind="S")
    return dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True), dce_raw


def test_vit_adapter_oracle_matches_reference():
    """use_backbone encoders (ViT-B/16 stand-in + BackboneAdapter necks + GroupNorm mix) and the fusion head on
    14 x 14 maps: the restatement against the unmodified reference modules' outputs (model_vit.npz)."""
    gold = gu.load("model_vit.npz")
    shapes = gu.load_shapes("vit")
    p, _ = vit_parameters()
    sds = {m: op.seeded_state_dict(shapes[m], seed=11) for m in ("dwi", "dce", "fusion")}
    dwi, dce = vit_inputs()
    torch.set_num_threads(8)
    with torch.no_grad():
        ld, ad, md = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
        lc, ac, mc = mo.encoder_forward(sds["dce"], "dce", p, dce)
        lf, mf, af = mo.fusion_forward(sds["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    n = 0
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            gu.check(gold, key, t, rtol=2e-5)
            n += 1
    assert n == 34


GEOMETRY_VARIANTS = {
    "cnn128": dict(input_size=128, shapes="cnn"),
    "cnn_s2": dict(downsample=(True, False, True), shapes="cnn"),
    "cnn_mf1": dict(mask_stage="f1", shapes="cnn_mf1"),
    "cnn_mf3": dict(mask_stage="f3", shapes="cnn_mf3"),
    "cnn_r2": dict(repeat_blocks=(2, 1, 2), shapes="cnn_r2"),
}


def variant_parameters(tag):
    """Non-default configurations: 128 x 128 ROIs (64 x 64 maps: strided mask-head stack, 2x2-averaging projector
    pool), a stride-2 block3 (16 x 16 f3; the fusion head takes the bilinear mask path), the mask head on f1 / f3,
    repeated bottlenecks.  Returns (parameters, input size, tag of the state-shape table)."""
    v = GEOMETRY_VARIANTS[tag]
    size = v.get("input_size", 64)
    p = pd.default_parameters(input_size=size)
    for m in ("dwi", "dce"):
        mp = p[f"{m}_model_parameters"]
        if "downsample" in v:
            mp["downsample"] = v["downsample"]
        if "mask_stage" in v:
            mp["mask_parameters"] = dict(mp["mask_parameters"], mask_stage=v["mask_stage"])
        if "repeat_blocks" in v:
            mp["repeat_blocks"] = v["repeat_blocks"]
    return p, size, v["shapes"]


def variant_inputs(size):
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, size=size, kind="S")
    return dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True), dce_raw


@pytest.mark.parametrize("tag", sorted(GEOMETRY_VARIANTS))
def test_geometry_variant_oracle_matches_reference(tag):
    gold = gu.load(f"model_{tag}.npz")
    p, size, shape_tag = variant_parameters(tag)
    shapes = gu.load_shapes(shape_tag)
    sds = {m: op.seeded_state_dict(shapes[m], seed=7) for m in ("dwi", "dce", "fusion")}
    dwi, dce = variant_inputs(size)
    torch.set_num_threads(8)
    with torch.no_grad():
        ld, ad, md = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
        lc, ac, mc = mo.encoder_forward(sds["dce"], "dce", p, dce)
        lf, mf, af = mo.fusion_forward(sds["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    n = 0
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            gu.check(gold, key, t, rtol=2e-5)
            n += 1
    assert n == 34


def resnet_parameters():
    """The reference's default backbone family (RadImageNet / resnet50 branches of build_medical_backbone)."""
    import foundation_model as fm

    p = pd.default_parameters(input_size=224)
    backbones = {}
    for m, c in (("dwi", 16), ("dce", 6)):
        mp = p[f"{m}_model_parameters"]
        mp["backbone_str"], mp["use_backbone"] = "radimagenet", True
        backbones[m] = fm.build_medical_backbone(p, None, m, in_channels=c)
    return p, backbones


def test_resnet_oracle_matches_torchvision_stand_in():
    """timm is absent (parity unpinned for the backbone itself): the restated ResNet-50 feature extractor against
    torchvision's ResNet dilated to output stride 8, same parameter names."""
    import torchvision
    from oracle import backbone_oracle as bo

    tv = torchvision.models.resnet50(replace_stride_with_dilation=[False, True, True])
    tv.conv1 = torch.nn.Conv2d(6, 64, 7, stride=2, padding=3, bias=False)
    shapes = {k: tuple(v.shape) for k, v in tv.state_dict().items() if not k.startswith("fc.")}
    sd = op.seeded_state_dict(shapes, seed=3)
    tv.load_state_dict(sd, strict=False)
    tv.eval()
    x = torch.rand(1, 6, 96, 96, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        t = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        ref = []
        for layer in (tv.layer1, tv.layer2, tv.layer3, tv.layer4):
            t = layer(t)
            ref.append(t)
        got = bo.resnet_features(sd, x)
    assert [tuple(f.shape) for f in got] == [(1, 256, 24, 24), (1, 512, 12, 12), (1, 1024, 12, 12), (1, 2048, 12, 12)]
    for a, b in zip(got, ref):
        assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item()


def test_resnet_adapter_oracle_matches_reference():
    gold = gu.load("model_resnet.npz")
    shapes = gu.load_shapes("resnet")
    p, _ = resnet_parameters()
    sds = {m: op.seeded_state_dict(shapes[m], seed=13) for m in ("dwi", "dce", "fusion")}
    dwi, dce = vit_inputs()
    torch.set_num_threads(8)
    with torch.no_grad():
        ld, ad, md = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
        lc, ac, mc = mo.encoder_forward(sds["dce"], "dce", p, dce)
        lf, mf, af = mo.fusion_forward(sds["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    n = 0
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            gu.check(gold, key, t, rtol=2e-5)
            n += 1
    assert n == 34


def test_dwi_normalize_oracle_matches_reference():
    gold = gu.load("normalizers.npz")
    dwi_raw, _, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    dwi_u, _, _, _ = op.synthetic_raw(3, seed=77, kind="U")
    gu.check(gold, "dwi/S", no.dwi_normalize_batch(dwi_raw[:3]), rtol=1e-6)
    gu.check(gold, "dwi/U", no.dwi_normalize_batch(dwi_u), rtol=1e-6)
    gu.check(gold, "dwi/E", no.dwi_normalize_batch(op.edge_cases()), rtol=1e-6)
    gu.check(gold, "dwi/U_noadc", no.dwi_normalize_batch(dwi_u, clip_z=(-2, 2.5), adc=False), rtol=1e-6)


def test_nyul_oracle_matches_reference():
    gold = gu.load("normalizers.npz")
    _, dce_raw, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    _, dce_u, _, _ = op.synthetic_raw(3, seed=77, kind="U")
    lm = no.nyul_fit(list(dce_raw[:8]), num_channels=6)
    assert np.array_equal(lm, gold["nyul/landmarks"])  # float64, bit-exact
    gu.check(gold, "nyul/S", no.nyul_transform_batch(dce_raw[8:], lm), rtol=1e-7)
    gu.check(gold, "nyul/U", no.nyul_transform_batch(dce_u, lm), rtol=1e-7)
    gu.check(gold, "nyul/ties", no.nyul_transform_batch(torch.round(dce_u * 20) / 20, lm), rtol=1e-7)


def test_adc_oracle_matches_reference():
    gold = gu.load("normalizers.npz")
    dwi_raw, _, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    adc = no.compute_adc_map(dwi_raw[0, :13], list(range(13)))
    assert np.allclose(adc.numpy(), gold["adc/map"], rtol=1e-6, atol=1e-9)


def test_percentile_rule_is_numpy_percentile():
    g = np.random.default_rng(0)
    for n in (4096, 50176, 17):
        x = g.random(n).astype(np.float32)
        ref = np.percentile(x, no.DEFAULT_LANDMARKS)
        assert np.array_equal(no.percentile_linear(np.sort(x), no.DEFAULT_LANDMARKS), ref)


def test_vit_oracle_matches_torchvision_stand_in():
    """timm is absent (parity unpinned); the restated ViT-B/16 feature extractor is cross-checked against
    torchvision's VisionTransformer - the same architecture - with the weights mapped across (small depth)."""
    from torchvision.models.vision_transformer import VisionTransformer

    from oracle import backbone_oracle as bo

    shapes = bo.vit_shapes(in_chans=6, depth=2)
    sd = op.seeded_state_dict(shapes, seed=3)
    tv = VisionTransformer(image_size=224, patch_size=16, num_layers=2, num_heads=12, hidden_dim=768, mlp_dim=3072)
    tv.conv_proj = torch.nn.Conv2d(6, 768, 16, 16)
    for blk in tv.encoder.layers:
        blk.ln_1.eps = blk.ln_2.eps = 1e-6
    tvsd = bo.to_torchvision(sd)
    tvsd["heads.head.weight"], tvsd["heads.head.bias"] = tv.heads.head.weight.data, tv.heads.head.bias.data
    tv.load_state_dict(tvsd)
    tv.eval()
    x = torch.rand(2, 6, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        feats = bo.vit_features(sd, x)
        t = tv._process_input(x)
        t = torch.cat([tv.class_token.expand(2, -1, -1), t], dim=1) + tv.encoder.pos_embedding
        for i, blk in enumerate(tv.encoder.layers):
            t = blk(t)
            ref = t[:, 1:].transpose(1, 2).reshape(2, 768, 14, 14)
            assert torch.allclose(feats[i], ref, rtol=1e-4, atol=1e-4), i


def test_resize_oracle_is_torchvision_resize():
    """a5: the oracle's resize against torchvision.transforms.Resize (what the reference calls), 64 -> 224."""
    from torchvision import transforms

    x = torch.rand(3, 6, 64, 64, generator=torch.Generator().manual_seed(5))
    ref = torch.stack([transforms.Resize(224)(c) for c in x])
    assert torch.equal(no.resize(x, 224), ref)
    # for an upsample the antialias filter degenerates to plain bilinear interpolation (what the kernel does)
    plain = torch.nn.functional.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False)
    assert torch.allclose(ref, plain, rtol=0, atol=1e-5)  # the separable antialias code path rounds differently
    # shrinking (ROIs larger than input_size): torchvision antialiases tensors, and so does the oracle - what
    # b200_resize_aa_c1 is held to; the plain bilinear taps would differ visibly here
    big = torch.rand(2, 6, 160, 160, generator=torch.Generator().manual_seed(6))
    ref_small = torch.stack([transforms.Resize(64)(c) for c in big])
    assert torch.equal(no.resize(big, 64), ref_small)
    plain_small = torch.nn.functional.interpolate(big, size=(64, 64), mode="bilinear", align_corners=False)
    assert (ref_small - plain_small).abs().max().item() > 1e-2


def test_oracle_on_trained_weights_matches_reference_pipeline():
    """The briefly trained fixture (tests/golden/trained_cnn.npz: int8 deltas on the seeded weights) and the
    reference pipeline's outputs on it (model_cnn_trained.npz): the oracle - numpy normalisers + functional
    networks - reproduces the first 8 of the 1 024 held-out cases, and the fixture itself is non-degenerate
    (every class predicted, top-2 margins reaching down to ~0)."""
    import json

    import parameters_default as pd

    gold = gu.load("model_cnn_trained.npz")
    hp = json.loads(str(gold["hp"]))
    sds = op.trained_state_dicts(gu.load("trained_cnn.npz"), gu.load_shapes("cnn"), seed=hp["weight_seed"])
    n = 8
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(hp["n_eval"], seed=hp["eval_seed"], kind="S")
    labels, _ = op.structured_targets(dwi_raw)
    assert np.array_equal(labels.numpy(), gold["labels"])
    lm = gold["landmarks"]
    dwi, dce = no.dwi_normalize_batch(dwi_raw[:n]), no.nyul_transform_batch(dce_raw[:n], lm)
    with torch.no_grad():
        lf, mf, af = mo.pipeline_forward(sds, pd.default_parameters(), dwi, dce)
    ref = torch.from_numpy(gold["fusion_logits"][:n])
    assert (lf - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert (af["gating_weights"] - torch.from_numpy(gold["gating"][:n])).abs().max().item() <= 1e-4
    ref_m = torch.from_numpy(gold["fusion_mask_sum"][:n])
    assert (mf.sum((1, 2, 3)) - ref_m).abs().max().item() <= 2e-4 * ref_m.abs().max().item()
    full = torch.from_numpy(gold["fusion_logits"])
    hist = torch.bincount(full.argmax(1), minlength=4)
    assert hist.min().item() >= 0.1 * hp["n_eval"], hist
    top = full.topk(2, dim=1).values
    assert (top[:, 0] - top[:, 1]).min().item() < 0.02
