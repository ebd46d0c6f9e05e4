"""Debug aid: where does the ViT-path error come from?  Runs the use_backbone encoders with the backbone's
features replaced by the fp32 oracle's (rounded once to bf16) and prints the error of every output against the
reference golden, next to the errors of the full bf16 product path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200path  # noqa: F401,E402
import golden_util as gu  # noqa: E402
import model_module as b_mm  # noqa: E402
from oracle import backbone_oracle as bo  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402
from oracle import params as op  # noqa: E402
from test_oracle_golden import vit_inputs, vit_parameters  # noqa: E402

DEV = "cuda"
gold = gu.load("model_vit.npz")
shapes = gu.load_shapes("vit")
p, backbones = vit_parameters()
mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p, backbones["dwi"]),
        "dce": b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"]), "fusion": b_mm.FusionModel(p)}
sds = {k: op.seeded_state_dict(shapes[k], seed=11) for k in mods}
for k, m in mods.items():
    m.load_state_dict(sds[k])
    m.to(DEV).eval()
dwi, dce = vit_inputs()
torch.set_num_threads(16)


def run(tag):
    with torch.no_grad():
        ld, ad, md = mods["dwi"](dwi.to(DEV))
        lc, ac, mc = mods["dce"](dce.to(DEV))
        lf, mf, af = mods["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
    torch.cuda.synchronize()
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac,
            "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    res = {}
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            res[key] = gu.check(gold, key, t, rtol=1e9)
    return res


full = run("full")
for m, x in (("dwi", dwi), ("dce", dce)):
    sd = sds[m]
    s = mo.SD(sd)
    with torch.no_grad():
        xg, _ = mo.se_block(x, s.sub("modality_attention"))
        pre = "backbone_adapter.backbone._orig_mod."  # loaded last, so these values win (shared module)
        feats = bo.vit_features({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, xg)
    chains = p[f"{m}_model_parameters"]["backbone_index_lists"]
    cats = [torch.cat([feats[j] for j in c], 1).permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16() for c in chains]
    with torch.no_grad():
        gate = mods[m](x.to(DEV))[1]["mod_attn_map"].view(x.shape[0], -1).contiguous()
        mine = mods[m].backbone._orig_mod.forward_chains(x.to(DEV), chains, gate)
    for a, b in zip(mine, cats):
        print(m, "chain buffer", tuple(a.shape), "err", ((a.float() - b.float()).abs().max() / b.float().abs().max()).item())
    mods[m].backbone._orig_mod.forward_chains = (lambda cats: (lambda x, ch, gate=None: cats))(cats)
exact = run("exact backbone")
print(f"{'output':38s} {'bf16 ViT':>10s} {'oracle ViT feats':>18s}")
for k in sorted(full, key=lambda k: -full[k]):
    print(f"{k:38s} {full[k]:10.2e} {exact[k]:18.2e}")
