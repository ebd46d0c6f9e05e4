"""Generates tests/golden/trained_cnn.npz and tests/golden/model_cnn_trained.npz from the UNMODIFIED reference
(authoring container only; needs /root/reference, read-only):

    python oracle/make_golden_trained.py

Why: with seeded random weights every structured case predicts the same class (per-class logit spread over cases
~1e-2 against a top-2 margin of ~0.5), so "argmax agreement" says nothing.  Here the
reference's own modules are TRAINED briefly on a structured synthetic task - AdamW steps of the single-modality
objective (composed as LightningSingleModel._shared_step does, code/train.py:294-400, with the reference's own loss
functions) for each encoder, then AdamW steps of the frozen-encoder fusion objective (code/train_fusion.py:203-296) -
so that the class histogram over 1 024 held-out cases is spread over all four classes and the top-2 margins reach down
to zero.  Stored:

* trained_cnn.npz - the trained weights as int8 deltas against oracle.params.seeded_state_dict(seed=7), one fp32 scale
  per tensor (tensors of <= 4 096 elements - BatchNorm statistics, biases - are stored exactly).  The FIXTURE weights are by definition `seeded + q * scale` (fp32, elementwise, reproducible anywhere);
  the reference outputs below are computed with exactly those, so the quantisation loses nothing.
* model_cnn_trained.npz - the unmodified reference pipeline (DWINormalize / NyulStandardizer -> encoders -> FusionModel,
  eval mode) on 1 024 held-out structured cases: all three logit blocks, gating weights, per-case sums of both
  encoder masks and of the fused mask, plus the Nyul landmarks the run was fitted with.

Nothing here is imported by the product.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import make_golden as mg  # noqa: E402  (also puts /root/reference/code on sys.path)
import make_golden_train as mgt  # noqa: E402
from oracle import params as op  # noqa: E402

HP = {"weight_seed": 7, "train_seed": 777, "n_train": 256, "batch": 32, "enc_steps": 96, "fusion_steps": 96,
      "lr": 2e-3, "weight_decay": 1e-2, "smoothing": 0.1, "gamma": 1.5, "eval_seed": 4321, "n_eval": 1024}


def reference_normalise(dwi_raw, dce_raw, nyul):
    import dataset as ref_dataset

    norm = ref_dataset.DWINormalize()
    dwi = torch.stack([norm(c) for c in dwi_raw])
    dce = torch.stack([nyul.transform(c, num_channels=dce_raw.shape[1]) for c in dce_raw])
    return dwi.float(), dce.float()


def train_encoder(mm, ref_loss, ref_train, p, method, model, x, masks, labels):
    mp = p[f"{method}_model_parameters"]
    lam = {"lambda_mask": mp["mask_parameters"]["lambda_mask"], "lambda_recon": mp["lambda_recon"],
           "lambda_mimic": mp["lambda_mimic"], "lambda_feat_norm": mp["lambda_feat_norm"]}
    smoother = ref_loss.LabelSmoothing(p["class_num"], HP["smoothing"])
    crit = ref_loss.SoftWeightedFocalLoss(HP["gamma"], torch.ones(p["class_num"]))
    dice = ref_loss.SoftDiceLoss()
    opt = torch.optim.AdamW(model.parameters(), lr=HP["lr"], weight_decay=HP["weight_decay"])
    model.train()
    nb = x.shape[0] // HP["batch"]
    for it in range(HP["enc_steps"]):
        sl = slice((it % nb) * HP["batch"], (it % nb + 1) * HP["batch"])
        xb, mb, yb = x[sl], masks[sl], labels[sl]
        opt.zero_grad(set_to_none=True)
        out, aux, mask_out = model(xb, mb)
        cls = crit(out, smoother(out, yb))
        ns = types.SimpleNamespace(device=xb.device, mimic_enabled=True, lambda_recon=lam["lambda_recon"],
                                   lambda_mimic=lam["lambda_mimic"])
        recon_w, mimic_w = ref_train.LightningSingleModel.compute_aux_losses(ns, aux, xb, aux["proj_pairs"], 1.0, True)
        total = (cls + lam["lambda_feat_norm"] * ref_train.compute_feat_norm_loss(aux, xb.device) +
                 lam["lambda_mask"] * dice(mask_out, mb) + lam["lambda_recon"] * recon_w + lam["lambda_mimic"] * mimic_w)
        total.backward()
        opt.step()
        acc = (out.argmax(1) == yb).float().mean().item()
        print(f"  {method} step {it:3d} loss {total.item():.4f} cls {cls.item():.4f} acc {acc:.2f}", flush=True)
    model.eval()


def train_fusion(ref_loss, ref_tf, ref_train, p, models, dwi, dce, masks, labels):
    smoother = ref_loss.LabelSmoothing(p["class_num"], HP["smoothing"])
    crit = ref_loss.SoftWeightedFocalLoss(HP["gamma"], torch.ones(p["class_num"]))
    dice = ref_loss.SoftDiceLoss()
    lam = {"lambda_mask": 0.2, "lambda_recon": 0.1, "lambda_mimic": 0.2}
    feats = []
    with torch.no_grad():  # frozen encoders (eval mode: the cached features are deterministic)
        for i in range(0, dwi.shape[0], HP["batch"]):
            _, ad, md = models["dwi"](dwi[i:i + HP["batch"]])
            _, ac, mc = models["dce"](dce[i:i + HP["batch"]])
            feats.append((ad["raw_feats"][-1], ac["raw_feats"][-1], md, mc))
    fm = models["fusion"]
    opt = torch.optim.AdamW(fm.parameters(), lr=HP["lr"], weight_decay=HP["weight_decay"])
    fm.train()
    nb = len(feats)
    for it in range(HP["fusion_steps"]):
        b = it % nb
        sl = slice(b * HP["batch"], (b + 1) * HP["batch"])
        f3d, f3c, md, mc = feats[b]
        opt.zero_grad(set_to_none=True)
        logits, fused_mask, aux = fm([f3d], [f3c], md, mc)
        cls = crit(logits, smoother(logits, labels[sl]))
        mask = ref_tf.safe_mask_loss(fused_mask, masks[sl], dice) / 3
        fused_input = torch.cat([dwi[sl], dce[sl]], dim=1)
        recon = ref_tf.compute_recon_list_loss(aux["recon_fused"], fused_input) / 3
        p1, p1_r, p2, p2_r = aux["proj_fused"][:4]
        mimic = (ref_train.mimic_feat_loss(p1, p1_r) + ref_train.mimic_feat_loss(p2, p2_r)) / 2
        total = cls + lam["lambda_mask"] * mask + lam["lambda_recon"] * recon + lam["lambda_mimic"] * mimic
        total.backward()
        opt.step()
        acc = (logits.argmax(1) == labels[sl]).float().mean().item()
        print(f"  fusion step {it:3d} loss {total.item():.4f} cls {cls.item():.4f} acc {acc:.2f}", flush=True)
    fm.eval()


def main():
    mgt.stub_harness_modules()
    import loss as ref_loss
    import model_module as mm
    import preprocess_helpers as ref_pre
    import train as ref_train
    import train_fusion as ref_tf

    t0 = time.time()
    p = mg.configure(mg.reference_parameters())
    torch.manual_seed(0)
    models = {"dwi": mm.ModelMaskHeadBackbone("dwi", p, None), "dce": mm.ModelMaskHeadBackbone("dce", p, None),
              "fusion": mm.FusionModel(p)}
    seeded = {}
    for name, m in models.items():
        seeded[name] = op.seeded_state_dict(op.shapes_of(m.state_dict()), seed=HP["weight_seed"])
        m.load_state_dict(seeded[name])

    dwi_raw, dce_raw, _, _ = op.synthetic_raw(HP["n_train"], seed=HP["train_seed"], kind="S")
    labels, masks = op.structured_targets(dwi_raw)
    print("train label histogram", torch.bincount(labels, minlength=4).tolist())
    nyul = ref_pre.NyulStandardizer()
    nyul.fit(list(dce_raw), num_channels=dce_raw.shape[1])
    dwi, dce = reference_normalise(dwi_raw, dce_raw, nyul)

    torch.manual_seed(1)
    train_encoder(mm, ref_loss, ref_train, p, "dwi", models["dwi"], dwi, masks, labels)
    train_encoder(mm, ref_loss, ref_train, p, "dce", models["dce"], dce, masks, labels)
    train_fusion(ref_loss, ref_tf, ref_train, p, models, dwi, dce, masks, labels)
    print(f"trained in {time.time() - t0:.0f} s")

    torch.save({k: m.state_dict() for k, m in models.items()}, "/tmp/trained_raw.pt")
    # --- quantise the update: fixture weights = seeded + q * scale -------------------------------------------
    out = {}
    for name, m in models.items():
        sd = m.state_dict()
        fixed = {}
        for k, v in sd.items():
            base = seeded[name][k]
            if not v.dtype.is_floating_point:
                fixed[k] = base.clone()
                continue
            if v.numel() <= 4096:
                # small tensors (BatchNorm affine / running statistics, biases, scalars) are stored exactly: a
                # quantised running variance could land at or below zero
                out[f"{name}/{k}/exact"] = v.detach().float().numpy().copy()
                fixed[k] = v.detach().clone()
                continue
            d = (v - base).float()
            scale = np.float32(d.abs().max().item() / 127.0)
            if scale == 0:
                fixed[k] = base.clone()
                continue
            q = torch.clamp(torch.round(d / float(scale)), -127, 127).to(torch.int8)
            out[f"{name}/{k}/q"] = q.numpy()
            out[f"{name}/{k}/scale"] = np.array(scale, dtype=np.float32)
            fixed[k] = op.apply_delta(base, q, scale)
        m.load_state_dict(fixed)
        m.eval()
    out["hp"] = np.array(json.dumps(HP))
    np.savez_compressed(os.path.join(mg.GOLD, "trained_cnn.npz"), **out)
    print("trained_cnn.npz", os.path.getsize(os.path.join(mg.GOLD, "trained_cnn.npz")) / 1e6, "MB")

    # --- reference outputs on the held-out cases --------------------------------------------------------------
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(HP["n_eval"], seed=HP["eval_seed"], kind="S")
    labels, _ = op.structured_targets(dwi_raw)
    dwi, dce = reference_normalise(dwi_raw, dce_raw, nyul)
    res = {k: [] for k in ("dwi_logits", "dce_logits", "fusion_logits", "gating", "dwi_mask_sum", "dce_mask_sum",
                           "fusion_mask_sum", "f3_dwi_sum", "f3_dce_sum")}
    with torch.no_grad():
        for i in range(0, HP["n_eval"], 64):
            ld, ad, md = models["dwi"](dwi[i:i + 64])
            lc, ac, mc = models["dce"](dce[i:i + 64])
            lf, mf, af = models["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
            res["dwi_logits"].append(ld), res["dce_logits"].append(lc), res["fusion_logits"].append(lf)
            res["gating"].append(af["gating_weights"])
            res["dwi_mask_sum"].append(md.sum((1, 2, 3))), res["dce_mask_sum"].append(mc.sum((1, 2, 3)))
            res["fusion_mask_sum"].append(mf.sum((1, 2, 3)))
            res["f3_dwi_sum"].append(ad["raw_feats"][-1].sum((1, 2, 3)))
            res["f3_dce_sum"].append(ac["raw_feats"][-1].sum((1, 2, 3)))
            print(f"  eval {i + 64}/{HP['n_eval']}", flush=True)
    gold = {k: torch.cat(v).numpy() for k, v in res.items()}
    gold["labels"] = labels.numpy()
    gold["landmarks"] = np.stack([nyul.channel_landmarks[c] for c in range(dce_raw.shape[1])])
    gold["dwi_norm_probe"] = dwi[::64, :, ::8, ::8].numpy()
    gold["dce_norm_probe"] = dce[::64, :, ::8, ::8].numpy()
    gold["hp"] = np.array(json.dumps(HP))
    np.savez_compressed(os.path.join(mg.GOLD, "model_cnn_trained.npz"), **gold)
    for k in ("dwi_logits", "dce_logits", "fusion_logits"):
        lg = torch.from_numpy(gold[k])
        top = lg.topk(2, dim=1).values
        margin = (top[:, 0] - top[:, 1])
        qs = torch.quantile(margin, torch.tensor([0.0, 0.01, 0.1, 0.5, 0.9]))
        print(k, "class histogram", torch.bincount(lg.argmax(1), minlength=4).tolist(), "accuracy",
              round((lg.argmax(1) == labels).float().mean().item(), 3), "max|logit|", round(lg.abs().max().item(), 3),
              "margin quantiles [0,1,10,50,90]%", [round(v, 4) for v in qs.tolist()])
    print(f"done in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
