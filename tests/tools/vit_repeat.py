"""Debug: the ViT-path encoders + fusion head run 4 times in one process (allocator poisoned with NaNs in between):
per output, the spread of the error against the reference fixture and the run-to-run difference."""
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
runs = []
for rep in range(4):
    junk = [torch.full((32 << 20,), float("nan"), device="cuda") for _ in range(8)]
    del junk
    (ld, ad, md), (lc, ac, mc), (lf, mf, af) = _run_product(mods, dwi, dce)
    outs = {"S/dwi/logits": ld, "S/dwi/aux": ad, "S/dwi/mask": md, "S/dce/logits": lc, "S/dce/aux": ac, "S/dce/mask": mc, "S/fusion/logits": lf, "S/fusion/mask": mf, "S/fusion/aux": af}
    flat = {}
    for prefix, obj in outs.items():
        for key, t in gu.walk(prefix, obj):
            flat[key] = (t.float().clone(), gu.check(gold, key, t, rtol=10.0))
    runs.append(flat)
for key in runs[0]:
    errs = [r[key][1] for r in runs]
    d = max(((r[key][0] - runs[0][key][0]).abs().max() / runs[0][key][0].abs().max().clamp_min(1e-12)).item() for r in runs[1:])
    nan = any(torch.isnan(r[key][0]).any().item() for r in runs)
    if max(errs) > 0.02 or d > 0 or nan:
        print(f"{key:32s} err vs fixture {min(errs):.4f} .. {max(errs):.4f}   run-to-run max diff {d:.2e}  nan={nan}")
