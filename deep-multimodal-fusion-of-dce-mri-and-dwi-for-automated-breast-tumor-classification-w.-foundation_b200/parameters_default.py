"""Configuration dict for the hot path, in the layout the reference modules read.

The reference builds one nested dict by executing code/parameters_generate.py (which also
writes to Google-Drive paths) and aliases the DWI, DCE and fusion sub-dicts to one object
(parameters_generate.py:174, :183).  This factory produces only the keys the path reads
(model_module.py:499-522, :831-851), with the reference's default values, and gives every
modality its own dict so per-modality edits do not leak.
"""
from __future__ import annotations

import copy


def default_model_parameters(input_size=64):
    return {
        "input_size": input_size,
        "use_hybrid_transformer": False,
        "transformer_heads": 4,
        "transformer_patch_size": 2,
        "transformer_depth": 6,
        "transformer_embed_dim": 512,
        "dropout": 0.2,
        "channels": (128, 256, 512),
        "repeat_blocks": (1, 1, 1),
        "downsample": (True, False, False),
        "downsample_each_repeat": False,
        "mid_squeeze": 2,
        "backbone_index_lists": [],
        "backbone_out_channels": (),
        "proj_dim": 64,
        "use_se": True,
        "enable_modality_attention": True,
        "use_backbone": False,
        "transformer_backbone": False,
        "backbone_str": "vit_base_patch16_224",
        "mask_parameters": {
            "mask": True,
            "mask_stage": "f2",
            "lambda_mask": 0.2,
            "mask_loss_type": "dice",
            "mask_target_size": (32, 32),
            "mask_fusion_attention": True,
        },
    }


def default_parameters(dwi_channels=16, dce_channels=6, input_size=64, class_num=4, batch_size=32):
    """BASELINE.json shapes: DWI 16 b-values, DCE 6 phases, 64x64 ROIs, 4 classes."""
    p = {
        "dim": 2,
        "class_num": class_num,
        "batch_size": batch_size,
        "namelist": ["train", "val", "test"],
        "methods": ["dwi", "dce"],
        "dwi_channel_num": dwi_channels,
        "dce_channel_num": dce_channels,
    }
    base = default_model_parameters(input_size)
    for m in ("dwi", "dce", "fusion"):
        p[f"{m}_model_parameters"] = copy.deepcopy(base)
    p["fusion_model_parameters"]["fusion_specific_parameters"] = {
        "mha_heads": 4,
        "use_cross_attention": True,
        "use_mask_attention": True,
        "token_pool": (4, 4),
        "fusion_channels": 128,
        "dwi_out_channels": base["channels"][-1],
        "dce_out_channels": base["channels"][-1],
        "fusion_recon_ch": 1,
    }
    return p
