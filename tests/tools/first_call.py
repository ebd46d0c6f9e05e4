"""Debug: poison the caching allocator with NaNs, then compare the first and second forward of the ViT-path encoders."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import b200path, torch
import golden_util as gu
import model_module as b_mm
from oracle import params as op
from test_oracle_golden import vit_inputs, vit_parameters
shapes = gu.load_shapes("vit")
p, backbones = vit_parameters()
m = b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"])
m.load_state_dict(op.seeded_state_dict(shapes["dce"], seed=11)); m.to("cuda").eval()
dwi, dce = vit_inputs()
x = dce.to("cuda")
poison = [torch.full((64 << 20,), float("nan"), device="cuda") for _ in range(24)]  # 6 GB of NaN blocks
sizes = [1 << k for k in range(10, 27)]
small = [torch.full((s,), float("nan"), device="cuda") for s in sizes for _ in range(8)]
del poison, small
outs = []
for rep in range(2):
    with torch.no_grad():
        l, a, mk = m(x)
    torch.cuda.synchronize()
    flat = {"logits": l, "mask": mk}
    for k, v in a.items():
        if isinstance(v, (list, tuple)):
            for i, t in enumerate(v): flat[f"{k}.{i}"] = t
        elif v is not None: flat[k] = v
    outs.append({k: v.float().clone() for k, v in flat.items()})
for k in outs[0]:
    a, b = outs[0][k], outs[1][k]
    print(f"{k:20s} nan1={int(torch.isnan(a).sum())} nan2={int(torch.isnan(b).sum())} maxdiff={float((a-b).abs().nan_to_num(9e9).max()):.4g} max={float(b.abs().max()):.4g}")
