"""B200-native drop-in for the reference's ``transformer_model`` (code/transformer_model.py).

Same class names, constructor signatures and parameter names as the reference's in-house
pre-norm transformer stage (PatchEmbed :7-32, TokensToFeatureMap :34-52, TransformerEncoder
:54-66, TransformerBlock :68-81, MultiHeadSelfAttention :83-116, MLP :118-134,
TransformerStage :137-175).  The modules are parameter containers; the stage's arithmetic is
scheduled by ModelMaskHeadBackbone.forward on the sm_100a kernels (linears through
b200_conv_gemm / b200_gemm_batched, LayerNorm through b200_layernorm; see
model_module._transformer_stage).  Calling a sub-module's forward on its own raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ["PatchEmbed", "TokensToFeatureMap", "TransformerEncoder", "TransformerBlock", "MultiHeadSelfAttention",
           "MLP", "TransformerStage"]


def _not_wired(name):
    raise NotImplementedError(f"{name} is a parameter container; the stage runs inside ModelMaskHeadBackbone.forward")


class PatchEmbed(nn.Module):
    def __init__(self, in_ch, embed_dim, patch_size=2, dim=2):
        super().__init__()
        if dim != 2:
            raise NotImplementedError("dim=2 only")
        self.dim = dim
        self.norm = nn.LayerNorm(embed_dim)
        self.proj = nn.Conv2d(in_ch, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        _not_wired("PatchEmbed")


class TokensToFeatureMap(nn.Module):
    def __init__(self, dim=2):
        super().__init__()
        self.dim = dim

    def forward(self, tokens, spatial_shape):
        B, N, C = tokens.shape
        H, W = spatial_shape
        return tokens.transpose(1, 2).reshape(B, C, H, W)  # pure view bookkeeping, no arithmetic


class MultiHeadSelfAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, qkv_bias=True, attn_drop=0.1, proj_drop=0.1):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(embed_dim, embed_dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        _not_wired("MultiHeadSelfAttention")


class MLP(nn.Module):
    def __init__(self, embed_dim, mlp_ratio=4.0, drop=0.1):
        super().__init__()
        hidden = int(embed_dim * mlp_ratio)
        self.fc1 = nn.Linear(embed_dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, embed_dim)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        _not_wired("MLP")


class TransformerBlock(nn.Module):
    def __init__(self, embed_dim, heads, init_scale=0.1):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.attn = MultiHeadSelfAttention(embed_dim, heads)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.mlp = MLP(embed_dim)
        self.gamma1 = nn.Parameter(init_scale * torch.ones(embed_dim))
        self.gamma2 = nn.Parameter(init_scale * torch.ones(embed_dim))

    def forward(self, x):
        _not_wired("TransformerBlock")


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, depth=4, heads=8):
        super().__init__()
        self.layers = nn.ModuleList([TransformerBlock(embed_dim, heads=heads) for _ in range(depth)])

    def forward(self, x):
        _not_wired("TransformerEncoder")


class TransformerStage(nn.Module):
    def __init__(self, in_ch, embed_dim, depth=2, heads=8, patch_size=2, dim=2):
        super().__init__()
        if dim != 2:
            raise NotImplementedError("dim=2 only")
        self.dim = dim
        self.patch_embed = PatchEmbed(in_ch=in_ch, embed_dim=embed_dim, patch_size=patch_size, dim=dim)
        self.transformer = TransformerEncoder(embed_dim=embed_dim, depth=depth, heads=heads)
        self.tokens_to_map = TokensToFeatureMap(dim=dim)

    def forward(self, x):
        _not_wired("TransformerStage")
