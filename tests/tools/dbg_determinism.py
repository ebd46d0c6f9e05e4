"""Debug aid: run the use_backbone encoders twice on the same input and report run-to-run differences."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200path  # noqa: F401,E402
import golden_util as gu  # noqa: E402
import model_module as b_mm  # noqa: E402
from oracle import params as op  # noqa: E402
from test_oracle_golden import vit_inputs, vit_parameters  # noqa: E402

DEV = "cuda"
shapes = gu.load_shapes("vit")
p, backbones = vit_parameters()
mods = {"dwi": b_mm.ModelMaskHeadBackbone("dwi", p, backbones["dwi"]),
        "dce": b_mm.ModelMaskHeadBackbone("dce", p, backbones["dce"])}
for k, m in mods.items():
    m.load_state_dict(op.seeded_state_dict(shapes[k], seed=11))
    m.to(DEV).eval()
dwi, dce = vit_inputs()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


for m, x in (("dwi", dwi), ("dce", dce)):
    xd = x.to(DEV)
    chains = p[f"{m}_model_parameters"]["backbone_index_lists"]
    with torch.no_grad():
        junk = torch.full((64 << 20,), float("nan"), device=DEV)  # poison freed memory between runs
        del junk
        a = [t.clone() for t in mods[m].backbone._orig_mod.forward_chains(xd, chains, None)]
        junk = torch.full((64 << 20,), float("nan"), device=DEV)
        del junk
        b = mods[m].backbone._orig_mod.forward_chains(xd, chains, None)
        print(m, "chains run-to-run:", [rel(u, v) for u, v in zip(a, b)])
        o1 = mods[m](xd)
        o1 = ([t.clone() for t in o1[1]["raw_feats"]], [t.clone() for t in o1[1]["recon_feats"]], o1[2].clone(), o1[0].clone())
        junk = torch.full((64 << 20,), float("nan"), device=DEV)
        del junk
        # dirty shared memory / TMEM / freed global memory with large finite values from unrelated launches
        import b200_native as nat
        for cin, cout, taps in ((64, 64, 9), (128, 128, 9), (256, 256, 9), (256, 512, 1), (128, 64, 1)):
            xx = (torch.randn(8, 32, 32, cin, device=DEV) * 1000).bfloat16()
            ww = torch.randn(cout, taps * cin, device=DEV).bfloat16()
            nat.conv_gemm(xx, ww, taps=taps, act=1)
        junk = torch.full((256 << 20,), 12345.0, device=DEV)
        del junk, xx, ww
        o2 = mods[m](xd)
        torch.cuda.synchronize()
        print(m, "raw_feats:", [rel(u, v) for u, v in zip(o1[0], o2[1]["raw_feats"])],
              "recon:", [rel(u, v) for u, v in zip(o1[1], o2[1]["recon_feats"])], "mask:", rel(o1[2], o2[2]),
              "logits:", rel(o1[3], o2[0]))
        print(m, "nan check:", [bool(torch.isnan(t.float()).any()) for t in o2[1]["raw_feats"]])
        print(m, "sums:", [t.double().sum().item() for t in o2[1]["raw_feats"]], o2[2].double().sum().item(),
              xd.double().sum().item(), [t.double().sum().item() for t in b])
