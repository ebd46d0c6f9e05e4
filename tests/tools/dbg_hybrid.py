import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import b200path, torch
import golden_util as gu
import test_parity_gpu as T
from oracle import params as op, model_oracle as mo
gold = gu.load("model_hybrid.npz")
p, sds, mods = T._build(hybrid=True)
for kind in ("U","S"):
    dwi_raw, dce_raw, _, _ = op.synthetic_raw(2, seed=1234, kind=kind)
    dwi = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    with torch.no_grad():
        ld, ad, md = mods["dwi"](dwi.cuda()); lc, ac, mc = mods["dce"](dce_raw.cuda())
        lf, mf, af = mods["fusion"](ad["raw_feats"], ac["raw_feats"], md, mc)
        # oracle fusion fed with OUR encoder outputs (isolates the fusion head)
        rd = [t.float().cpu() for t in ad["raw_feats"]]; rc = [t.float().cpu() for t in ac["raw_feats"]]
        of = mo.fusion_forward(sds["fusion"], p, rd, rc, md.cpu(), mc.cpu())
    def rel(a,b): 
        a=a.float().cpu(); b=b.float().cpu(); return ((a-b).abs().max()/b.abs().max()).item()
    print(kind, "fusion-head-only errors vs oracle on same inputs: logits", rel(lf, of[0]), "mask", rel(mf, of[1]),
          {k: round(rel(af[k], of[2][k]),5) for k in af})
    for prefix, obj in {f"{kind}/dwi/aux": ad, f"{kind}/dce/aux": ac, f"{kind}/fusion/logits": lf, f"{kind}/fusion/mask": mf, f"{kind}/fusion/aux": af}.items():
        for key, t in gu.walk(prefix, obj):
            try:
                e = gu.check(gold, key, t, rtol=10.0)
            except AssertionError as ex:
                e = str(ex)[:80]
            if "raw_feats" in key or "fusion" in key: print("  ", key, e)
