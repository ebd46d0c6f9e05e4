import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import b200path, torch
import golden_util as gu
import model_module as b_mm
from oracle import params as op
from test_oracle_golden import vit_inputs, vit_parameters
from test_parity_gpu import _run_product
gold = gu.load("model_vit.npz"); shapes = gu.load_shapes("vit")
p, backbones = vit_parameters()
mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p, backbones["dwi"]), "dce": b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"]), "fusion": b_mm.FusionModel(p)}
for k, m in mods.items():
    m.load_state_dict(op.seeded_state_dict(shapes[k], seed=11)); m.to("cuda").eval()
dwi, dce = vit_inputs()
(ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac, "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
worst = {}
for prefix, obj in outs.items():
    for key, t in gu.walk(prefix, obj):
        worst[key] = gu.check(gold, key, t, rtol=10.0)
for k, v in sorted(worst.items(), key=lambda kv: -kv[1]): print(f"{v:.4f} {k}")
import numpy as np
for rep in range(3):
    (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
    r2 = ac["recon_feats"][1].float().cpu()
    ref = torch.from_numpy(gold["S/dce/aux.recon_feats.1/full"]) if "S/dce/aux.recon_feats.1/full" in gold else None
    print("r2 shape", tuple(r2.shape), "keys", [k for k in gold.keys() if "dce/aux.recon_feats.1" in k])
    if ref is not None:
        d = (r2 - ref).abs()
        idx = np.unravel_index(d.argmax().item(), d.shape)
        print(rep, "max diff", d.max().item(), "at", idx, "ref max", ref.abs().max().item(), "n>0.05*max", (d > 0.05 * ref.abs().max()).sum().item())
