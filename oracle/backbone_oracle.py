"""fp32 CPU oracle of the ViT-B/16 `features_only` backbone (test infrastructure only).

The reference obtains this backbone from the third-party package timm, which is neither vendored
nor pinned and is absent here (no network): /root/reference/code/foundation_model.py:371-431
(`timm.create_model("vit_base_patch16_224", features_only=True, out_indices=0..11, img_size=...,
in_chans=C)`), dispatched from :526-545.  PARITY UNPINNED: there is no reference output to pin to.
This file restates the published algorithm of timm's VisionTransformer as the reference uses it
(SURVEY.md row a13): Conv2d(C,768,16,16) patch embedding, cls token + learned position embedding,
12 pre-norm blocks (LayerNorm eps 1e-6, fused qkv with bias, 12 heads x 64, MLP 3072, exact GELU, no
LayerScale), every block's patch tokens (cls stripped, no final norm) reshaped to [B,768,14,14].
tests/test_oracle_golden.py cross-checks it against torchvision's VisionTransformer (same architecture,
present in this image) with the weights mapped across.

State-dict keys follow timm: patch_embed.proj.{weight,bias}, cls_token, pos_embed,
blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.{weight,bias}.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LN_EPS = 1e-6


def vit_features(sd, x, patch=16, heads=12):
    """x [B,C,H,W] fp32 -> list of per-block feature maps [B,E,H/patch,W/patch]."""
    b = x.shape[0]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    e, gh, gw = t.shape[1], t.shape[2], t.shape[3]
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(b, -1, -1), t], dim=1) + sd["pos_embed"]
    dh = e // heads
    feats = []
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p = f"blocks.{i}."
        h = F.layer_norm(t, (e,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        n = qkv.shape[1]
        qkv = qkv.reshape(b, n, 3, heads, dh).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * dh ** -0.5).softmax(dim=-1)
        h = (attn @ qkv[2]).transpose(1, 2).reshape(b, n, e)
        t = t + F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F.layer_norm(t, (e,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
        h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        t = t + h
        feats.append(t[:, 1:].transpose(1, 2).reshape(b, e, gh, gw))
        i += 1
    return feats


def vit_shapes(in_chans, img=224, patch=16, embed=768, depth=12, mlp=3072):
    """Parameter shapes of the backbone (timm key names)."""
    n = (img // patch) ** 2 + 1
    s = {"patch_embed.proj.weight": (embed, in_chans, patch, patch), "patch_embed.proj.bias": (embed,),
         "cls_token": (1, 1, embed), "pos_embed": (1, n, embed), "norm.weight": (embed,), "norm.bias": (embed,)}
    for i in range(depth):
        p = f"blocks.{i}."
        s.update({p + "norm1.weight": (embed,), p + "norm1.bias": (embed,), p + "attn.qkv.weight": (3 * embed, embed),
                  p + "attn.qkv.bias": (3 * embed,), p + "attn.proj.weight": (embed, embed), p + "attn.proj.bias": (embed,),
                  p + "norm2.weight": (embed,), p + "norm2.bias": (embed,), p + "mlp.fc1.weight": (mlp, embed),
                  p + "mlp.fc1.bias": (mlp,), p + "mlp.fc2.weight": (embed, mlp), p + "mlp.fc2.bias": (embed,)})
    return s


def to_torchvision(sd):
    """Map timm-style keys onto torchvision.models.vision_transformer.VisionTransformer's state dict."""
    out = {"conv_proj.weight": sd["patch_embed.proj.weight"], "conv_proj.bias": sd["patch_embed.proj.bias"],
           "class_token": sd["cls_token"], "encoder.pos_embedding": sd["pos_embed"],
           "encoder.ln.weight": sd["norm.weight"], "encoder.ln.bias": sd["norm.bias"]}
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p, q = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        out.update({q + "ln_1.weight": sd[p + "norm1.weight"], q + "ln_1.bias": sd[p + "norm1.bias"],
                    q + "self_attention.in_proj_weight": sd[p + "attn.qkv.weight"],
                    q + "self_attention.in_proj_bias": sd[p + "attn.qkv.bias"],
                    q + "self_attention.out_proj.weight": sd[p + "attn.proj.weight"],
                    q + "self_attention.out_proj.bias": sd[p + "attn.proj.bias"],
                    q + "ln_2.weight": sd[p + "norm2.weight"], q + "ln_2.bias": sd[p + "norm2.bias"],
                    q + "mlp.0.weight": sd[p + "mlp.fc1.weight"], q + "mlp.0.bias": sd[p + "mlp.fc1.bias"],
                    q + "mlp.3.weight": sd[p + "mlp.fc2.weight"], q + "mlp.3.bias": sd[p + "mlp.fc2.bias"]})
        i += 1
    return out


# ------------------------------------------------------------------------- ResNet-50 ----
def resnet_features(sd, x, layers=(3, 4, 6, 3), output_stride=8):
    """fp32 restatement of timm / torchvision `resnet50` (v1.5) as a feature extractor - what the reference builds
    at /root/reference/code/foundation_model.py:15-68 and :243-250 (`timm.create_model("resnet50",
    features_only=True, output_stride=8, out_indices=(1, 2, 3, 4), in_chans=C)`): returns [C2, C3, C4, C5].
    timm is absent (PARITY UNPINNED for the backbone itself); tests/test_oracle_golden.py cross-checks this
    restatement against torchvision's ResNet with `replace_stride_with_dilation`, whose parameter names it shares."""
    def bn(t, p):
        return F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.0, 1e-5)

    t = F.relu(bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), "bn1"))
    t = F.max_pool2d(t, 3, stride=2, padding=1)
    feats = []
    net_stride, dilation, prev_dilation = 4, 1, 1
    for li, (n_blocks, stride) in enumerate(zip(layers, (1, 2, 2, 2))):
        if net_stride >= output_stride:
            dilation *= stride
            stride = 1
        else:
            net_stride *= stride
        for bi in range(n_blocks):
            p = f"layer{li + 1}.{bi}."
            s, d = (stride, prev_dilation) if bi == 0 else (1, dilation)
            h = F.relu(bn(F.conv2d(t, sd[p + "conv1.weight"]), p + "bn1"))
            h = F.relu(bn(F.conv2d(h, sd[p + "conv2.weight"], stride=s, padding=d, dilation=d), p + "bn2"))
            h = bn(F.conv2d(h, sd[p + "conv3.weight"]), p + "bn3")
            idn = t
            if p + "downsample.0.weight" in sd:
                idn = bn(F.conv2d(t, sd[p + "downsample.0.weight"], stride=s), p + "downsample.1")
            t = F.relu(h + idn)
        prev_dilation = dilation
        feats.append(t)
    return feats
