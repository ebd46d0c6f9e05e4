"""fp32 CPU oracle of the encoders and the fusion head (test infrastructure only).

A functional restatement over a flat state dict (same keys as the reference modules'
``state_dict()``) of, in eval mode (BatchNorm uses running statistics, Dropout is identity):
  * ModelMaskHeadBackbone.forward   /root/reference/code/model_module.py:645-733
  * ResNetLiteBlock_withRecon       /root/reference/code/model_module.py:298-316
  * SEBlock / ReconHead / MaskHeadResize / MaskGuidedSpatialAttention / Projector /
    ClassificationHead / FeatureDownAlign  model_module.py:25-396
  * TransformerStage and parts      /root/reference/code/transformer_model.py:7-175
  * FusionModel.forward             /root/reference/code/model_module.py:919-1000
It is validated against the unmodified reference modules by oracle/make_golden.py and
against the committed fixtures by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default


class SD:
    """Prefix view over a flat state dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return SD(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


BN_BATCH_STATS = False  # train-mode BatchNorm (batch statistics); set by oracle.train_oracle around its forward only


def _bn(x, s: SD):
    if BN_BATCH_STATS:  # nn.BatchNorm2d in train mode: normalise with the batch statistics (running stats untouched here)
        return F.batch_norm(x, None, None, s["weight"], s["bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, s["running_mean"], s["running_var"], s["weight"], s["bias"], False, 0.0, BN_EPS)


def _conv(x, s: SD, stride=1, padding=0):
    return F.conv2d(x, s["weight"], s["bias"] if s.has("bias") else None, stride=stride, padding=padding)


def se_block(x, s: SD):
    """model_module.py:34-43."""
    w = x.mean(dim=(2, 3), keepdim=True)
    w = torch.sigmoid(_conv(F.gelu(_conv(w, s.sub("fc.1"))), s.sub("fc.3")))
    return x * w, w


def recon_head(x, s: SD):
    """model_module.py:113-125 (upsample=False)."""
    y = F.gelu(_bn(_conv(x, s.sub("conv.0"), padding=1), s.sub("conv.1")))
    return _conv(y, s.sub("conv.3"), padding=1)


def res_block(x, s: SD, stride, num_repeats, downsample_each_repeat, use_se, drop_p=0.0):
    """model_module.py:298-316.  drop_p > 0: the block's nn.Dropout layers active (MC-dropout inference,
    train_fusion.py:445-481) with torch's global generator; BatchNorm stays in eval mode."""
    identity = _bn(_conv(x, s.sub("skip.0"), stride=stride), s.sub("skip.1")) if s.has("skip.0.weight") else x
    out = x
    for i in range(num_repeats):
        b = s.sub(f"bottlenecks.{i}")
        st = stride if (i == 0 or downsample_each_repeat) else 1
        out = F.gelu(_bn(_conv(out, b.sub("0"), stride=st), b.sub("1")))
        out = F.dropout(out, drop_p, training=drop_p > 0)                      # :260
        out = F.gelu(_bn(_conv(out, b.sub("4"), padding=1), b.sub("5")))
        out = _bn(_conv(out, b.sub("7")), b.sub("8"))
    out = F.gelu(out + identity)
    out = F.dropout(out, drop_p, training=drop_p > 0)                          # :305-306
    if use_se:
        out, _ = se_block(out, s.sub("se"))
    rec = recon_head(out, s.sub("reconstruct")) if s.has("reconstruct.conv.0.weight") else None
    return out, rec


def mask_head(x, s: SD, out_size=32):
    """model_module.py:197-215."""
    x = _conv(x, s.sub("pre"))
    size = x.shape[-1]
    table = {64: ("down_64_to_32", 1), 128: ("down_128_to_32", 2), 256: ("down_256_to_32", 3),
             512: ("down_512_to_32", 4)}
    if size == 32:
        pass
    elif size in table:
        name, n = table[size]
        for i in range(n):
            x = F.gelu(_conv(x, s.sub(f"{name}.{2 * i}"), stride=2, padding=1))
    else:
        x = F.interpolate(x, size=(out_size, out_size), mode="bilinear", align_corners=False)
    return _conv(x, s.sub("out"))


def mask_spatial_attention(img, mask, s: SD):
    """model_module.py:75-97."""
    if mask.shape[-2:] != img.shape[-2:]:
        mask = F.interpolate(mask, size=img.shape[-2:], mode="bilinear", align_corners=False)
    p = s.sub("mask_processor")
    a = F.conv2d(mask, p["0.weight"])
    a = F.group_norm(a, 1, p["1.weight"], p["1.bias"], 1e-5)
    a = torch.sigmoid(_conv(F.gelu(a), p.sub("3")))
    a = torch.clamp(a, 1e-4, 1.0 - 1e-4)
    return img * (1 + s["gamma"] * a), a


def projector(x, s: SD):
    """model_module.py:337-348."""
    x = F.gelu(_bn(_conv(x, s.sub("proj.0")), s.sub("proj.1")))
    return F.gelu(_bn(_conv(x, s.sub("proj.3")), s.sub("proj.4")))


def classification_head(x, s: SD, normalize=True):
    """model_module.py:364-369."""
    v = x.mean(dim=(2, 3))
    if normalize:
        v = F.normalize(v, dim=1)
    return F.linear(v, s["fc.weight"], s["fc.bias"])


def feature_down_align(x, s: SD):
    """model_module.py:386-396 with downsample=False (1x1 conv + BN + GELU, or identity)."""
    if not s.has("proj.0.weight"):
        return x
    return F.gelu(_bn(_conv(x, s.sub("proj.0")), s.sub("proj.1")))


# ------------------------------------------------------------ transformer stage --
def transformer_stage(x, s: SD, heads, patch):
    """transformer_model.py:137-175 (eval: every Dropout is identity)."""
    pe = s.sub("patch_embed")
    t = F.conv2d(x, pe["proj.weight"], pe["proj.bias"], stride=patch)  # :25
    hs, ws = t.shape[-2:]
    t = t.flatten(2).transpose(1, 2)                                    # :28
    e = t.shape[-1]
    t = F.layer_norm(t, (e,), pe["norm.weight"], pe["norm.bias"])       # :29
    i = 0
    while s.has(f"transformer.layers.{i}.norm1.weight"):
        l = s.sub(f"transformer.layers.{i}")
        h = F.layer_norm(t, (e,), l["norm1.weight"], l["norm1.bias"])
        t = t + _mhsa(h, l.sub("attn"), heads) * l["gamma1"]           # :79
        h = F.layer_norm(t, (e,), l["norm2.weight"], l["norm2.bias"])
        h = F.linear(F.gelu(F.linear(h, l["mlp.fc1.weight"], l["mlp.fc1.bias"])), l["mlp.fc2.weight"],
                     l["mlp.fc2.bias"])                                 # :128-133
        t = t + h * l["gamma2"]                                         # :80
        i += 1
    return t.transpose(1, 2).reshape(x.shape[0], e, hs, ws)            # :52


def _mhsa(x, s: SD, heads):
    """transformer_model.py:98-116."""
    b, n, c = x.shape
    dh = c // heads
    qkv = F.linear(x, s["qkv.weight"], s["qkv.bias"]).reshape(b, n, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * dh ** -0.5).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(y, s["proj.weight"], s["proj.bias"])


# ----------------------------------------------------------- backbone adapter ----
def backbone_adapter(x, s: SD, chains, feats=None):
    """BackboneAdapter.forward, model_module.py:452-476: backbone features -> per chain channel concat ->
    neck (3x3 conv + BN + GELU, twice, :440-447).  The backbone is the ViT-B/16 restated in
    oracle/backbone_oracle.py (its weights sit under `backbone.`); `feats` overrides it."""
    if feats is None:
        from oracle.backbone_oracle import resnet_features, vit_features
        pre = s.prefix + "backbone."
        if any(k.startswith(pre + "_orig_mod.") for k in s.sd):  # torch._dynamo.disable wrapper, model_module.py:539
            pre += "_orig_mod."
        bsd = {k[len(pre):]: v for k, v in s.sd.items() if k.startswith(pre)}
        feats = resnet_features(bsd, x) if "layer1.0.conv1.weight" in bsd else vit_features(bsd, x)
    outs = []
    for i, chain in enumerate(chains):
        n = s.sub(f"necks.f{i + 1}")
        y = torch.cat([feats[j] for j in chain], dim=1)                        # :471
        y = F.gelu(_bn(_conv(y, n.sub("0"), padding=1), n.sub("1")))
        outs.append(F.gelu(_bn(_conv(y, n.sub("3"), padding=1), n.sub("4"))))
    return outs


# -------------------------------------------------------------------- encoder ----
def encoder_forward(sd, method, params, x, backbone_feats=None, mc_dropout=False):
    """ModelMaskHeadBackbone.forward, model_module.py:645-733 (use_backbone: the ViT-B/16 adapter path).

    Returns (logits, aux, mask_pred) with the reference's aux keys.
    """
    s = SD(sd)
    mp = params[f"{method}_model_parameters"]
    use_se = mp["use_se"]
    reps = mp["repeat_blocks"]
    der = mp["downsample_each_repeat"]
    strides = [2 if d else 1 for d in mp["downsample"]]
    maskp = mp["mask_parameters"]
    mask_on, stage = maskp["mask"], maskp["mask_stage"].lower()
    size = maskp["mask_target_size"][0]
    dp = float(mp["dropout"]) if mc_dropout else 0.0

    mod_attn = None
    if mp["enable_modality_attention"]:
        x, mod_attn = se_block(x, s.sub("modality_attention"))                 # :649-650
    f2_b = f3_b = None
    if mp["use_backbone"]:
        x, f2_b, f3_b = backbone_adapter(x, s.sub("backbone_adapter"), mp["backbone_index_lists"], backbone_feats)
    f1, r1 = res_block(x, s.sub("block1"), strides[0], reps[0], der, use_se, dp)   # :666
    mask_pred = attn_map = None
    if mask_on and stage == "f1":
        mask_pred = mask_head(f1, s.sub("mask_head"), size)
        f1, attn_map = mask_spatial_attention(f1, mask_pred, s.sub("mask_spatial_attention"))
    f2_in = f1
    if mp["use_backbone"]:                                                      # :673-675
        a = torch.sigmoid(s["f2_weight"])
        f2_in = F.group_norm(a * f2_b + (1 - a) * f1, f1.shape[1], s["norm_f2.weight"], s["norm_f2.bias"], 1e-5)
    f2, r2 = res_block(f2_in, s.sub("block2"), strides[1], reps[1], der, use_se, dp)  # :679
    if mask_on and stage == "f2":
        m_in = f2 + feature_down_align(f1, s.sub("f1_to_f2"))                  # :682-683
        mask_pred = mask_head(m_in, s.sub("mask_head"), size)                  # :684
        f2, attn_map = mask_spatial_attention(f2, mask_pred, s.sub("mask_spatial_attention"))
    if not mp["use_hybrid_transformer"]:
        f3_in = f2
        if mp["use_backbone"]:                                                  # :688-690
            a = torch.sigmoid(s["f3_weight"])
            f3_in = F.group_norm(a * f3_b + (1 - a) * f2, f2.shape[1], s["norm_f3.weight"], s["norm_f3.bias"], 1e-5)
        f3, _ = res_block(f3_in, s.sub("block3"), strides[2], reps[2], der, use_se, dp)   # :694
        if mask_on and stage == "f3":
            m_in = f3 + feature_down_align(f2, s.sub("f2_to_f3"))
            mask_pred = mask_head(m_in, s.sub("mask_head"), size)
            f3, attn_map = mask_spatial_attention(f3, mask_pred, s.sub("mask_spatial_attention"))
    else:
        mid = transformer_stage(f2, s.sub("transformer"), mp["transformer_heads"], mp["transformer_patch_size"])
        f3 = _conv(mid, s.sub("trans_out_proj"))                               # :702-703
    pd = mp["proj_dim"]
    pool = lambda t: F.adaptive_avg_pool2d(t, (pd, pd))                         # :534, :707-710
    p1 = projector(pool(f1), s.sub("proj_f1"))
    p2 = projector(pool(f2), s.sub("proj_f2"))
    p1r = projector(pool(r1), s.sub("proj_r1"))
    p2r = projector(pool(r2), s.sub("proj_r2"))
    logits = classification_head(f3, s.sub("classification_head"))             # :720
    aux = {"raw_feats": [f1, f2, f3], "recon_feats": [r1, r2], "proj_pairs": [p1, p1r, p2, p2r],
           "mask_attn_map": attn_map, "mod_attn_map": mod_attn}
    return logits, aux, mask_pred


# --------------------------------------------------------------------- fusion ----
def fusion_forward(sd, params, raw_dwi, raw_dce, dwi_mask=None, dce_mask=None, run_dead_branch=False):
    """FusionModel.forward, model_module.py:919-1000.

    The cat -> fusion_conv_reduce -> refine branch (:935-940) never reaches an output
    (SURVEY.md appendix A-1); it is evaluated only when run_dead_branch is set (used when
    this oracle is timed as the CPU baseline, to do the work the reference does).
    """
    s = SD(sd)
    fc = params["fusion_model_parameters"]
    fs = fc["fusion_specific_parameters"]
    f3d, f3c = raw_dwi[-1], raw_dce[-1]
    p_dwi = F.conv2d(f3d, s["proj_in_dwi.weight"]) if s.has("proj_in_dwi.weight") else f3d     # :930
    p_dce = F.conv2d(f3c, s["proj_in_dce.weight"]) if s.has("proj_in_dce.weight") else f3c     # :931
    if run_dead_branch:
        red = F.gelu(_bn(_conv(torch.cat([p_dwi, p_dce], 1), s.sub("fusion_conv_reduce.reduce.0")),
                         s.sub("fusion_conv_reduce.reduce.1")))
        resid, _ = res_block(red, s.sub("refine"), 1, 1, False, False)
        F.gelu(red + resid)
    pv_d, pv_c = p_dwi.mean(dim=(2, 3)), p_dce.mean(dim=(2, 3))                                  # :948-949
    if fs["use_mask_attention"] and dwi_mask is not None and dce_mask is not None:              # :763-775
        xg = torch.cat([pv_d, pv_c, dwi_mask.mean(dim=(2, 3)), dce_mask.mean(dim=(2, 3))], 1)
    else:
        xg = torch.cat([pv_d, pv_c], 1)
    gw = torch.softmax(F.linear(xg, s["gating.fc.weight"], s["gating.fc.bias"]), dim=1)         # :779
    fused = gw[:, 0].view(-1, 1, 1, 1) * p_dwi + gw[:, 1].view(-1, 1, 1, 1) * p_dce              # :958
    attn_w = None
    if fs["use_cross_attention"]:
        hp, wp = fs["token_pool"]
        c = p_dwi.shape[1]
        tok = lambda t: F.adaptive_avg_pool2d(t, (hp, wp)).flatten(2).permute(0, 2, 1)          # :914-916
        ca = s.sub("cross_attn_block")
        a_out, attn_w = F.multi_head_attention_forward(                                          # :816
            tok(p_dwi).transpose(0, 1), tok(p_dce).transpose(0, 1), tok(p_dce).transpose(0, 1), c, fs["mha_heads"],
            ca["cross_attn.in_proj_weight"], ca["cross_attn.in_proj_bias"], None, None, False, 0.0,
            ca["cross_attn.out_proj.weight"], ca["cross_attn.out_proj.bias"], training=False, need_weights=True,
            average_attn_weights=True)
        a_out = a_out.transpose(0, 1)
        h = F.layer_norm(a_out, (c,), ca["attn_ffn.0.weight"], ca["attn_ffn.0.bias"])
        h = F.linear(F.gelu(F.linear(h, ca["attn_ffn.1.weight"], ca["attn_ffn.1.bias"])), ca["attn_ffn.3.weight"],
                     ca["attn_ffn.3.bias"])
        a_out = a_out + h                                                                        # :817
        low = a_out.permute(0, 2, 1).reshape(-1, c, hp, wp)                                      # :970
        fused = fused + F.interpolate(low, size=fused.shape[-2:], mode="bilinear", align_corners=False)  # :972-973
    if fc["use_se"]:
        fused, _ = se_block(fused, s.sub("fusion_se"))                                           # :977-978
    size = fc["mask_parameters"]["mask_target_size"][0]
    mask_logits = mask_head(fused, s.sub("mask_head"), size)                                     # :983
    logits = F.linear(fused.mean(dim=(2, 3)), s["classifier.2.weight"], s["classifier.2.bias"])  # :986
    recon = recon_head(fused, s.sub("fusion_reconstruct"))                                       # :989
    proj = projector(fused, s.sub("projF"))                                                      # :990
    aux = {"proj_fused": proj, "recon_fused": recon, "gating_weights": gw, "attn_weights": attn_w,
           "p_dwi": p_dwi, "p_dce": p_dce}
    return logits, mask_logits, aux


def pipeline_forward(sds, params, dwi, dce, run_dead_branch=False):
    """model_test.py:135-147: both encoders, then the fusion model on their raw features."""
    _, aux_d, m_d = encoder_forward(sds["dwi"], "dwi", params, dwi)
    _, aux_c, m_c = encoder_forward(sds["dce"], "dce", params, dce)
    return fusion_forward(sds["fusion"], params, aux_d["raw_feats"], aux_c["raw_feats"], m_d, m_c, run_dead_branch)
