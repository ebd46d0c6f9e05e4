"""How much error do bf16 GEMM operands alone introduce on the ViT-adapter path (CPU, oracle only)?

Runs the fp32 oracle twice on the tests/golden/model_vit.npz configuration - once as is, once with every
GEMM-shaped op (Conv2d / Linear with >= 64 input channels) fed bf16-rounded activations and weights (fp32
accumulation, nothing else changed) - and prints max|d| / max|ref| per API output.  With the seeded random
weights the two GroupNorm(C, C) backbone mixes (instance norms over 196 pixels) and the heavy-tailed feature
maps they produce amplify bf16 operand rounding to 2-7 % on the downstream maps; this is the floor the
tolerance of tests/test_parity_gpu.py::test_vit_adapter_pipeline_vs_golden_reference is set against.

    python tests/tools/bf16_floor.py            # ViT-adapter fixture (model_vit.npz configuration)
    python tests/tools/bf16_floor.py resnet     # ResNet-50 fixture (model_resnet.npz configuration)
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200path, golden_util as gu
from oracle import model_oracle as mo, params as op, backbone_oracle as bo
from test_oracle_golden import resnet_parameters, vit_inputs, vit_parameters
torch.set_num_threads(8)
RESNET = len(sys.argv) > 1 and sys.argv[1] == "resnet"
shapes = gu.load_shapes("resnet" if RESNET else "vit"); p, _ = resnet_parameters() if RESNET else vit_parameters()
SEED = 13 if RESNET else 11
dwi, dce = vit_inputs()
bf = lambda t: t.bfloat16().float()
class Emu:
    """F stand-in that rounds GEMM operands (activations + weights) to bf16, fp32 accumulate."""
    def __getattr__(self, n): return getattr(F, n)
    def conv2d(self, x, w, b=None, **k):
        if w.shape[1] >= 64: x, w = bf(x), bf(w)
        return F.conv2d(x, w, b, **k)
    def linear(self, x, w, b=None):
        if w.shape[1] >= 64: x, w = bf(x), bf(w)
        return F.linear(x, w, b)
def run(emu):
    mo.F = Emu() if emu else F; bo.F = Emu() if emu else F
    out = {}
    with torch.no_grad():
        for m, x in (("dwi", dwi), ("dce", dce)):
            sd = op.seeded_state_dict(shapes[m], seed=SEED)
            out[m] = mo.encoder_forward(sd, m, p, x)
        sdf = op.seeded_state_dict(shapes["fusion"], seed=SEED)
        out["fusion"] = mo.fusion_forward(sdf, p, out["dwi"][1]["raw_feats"], out["dce"][1]["raw_feats"], out["dwi"][2], out["dce"][2])
    return out
a = run(False); b = run(True)
def rel(x, y): return ((x-y).abs().max()/y.abs().max()).item()
for m in ("dwi", "dce"):
    print(m, "logits", rel(b[m][0], a[m][0]), "mask", rel(b[m][2], a[m][2]))
    for k, v in a[m][1].items():
        if isinstance(v, list): print(m, k, [rel(q, r) for q, r in zip(b[m][1][k], v)])
        elif v is not None: print(m, k, rel(b[m][1][k], v))
print("fusion logits", rel(b["fusion"][0], a["fusion"][0]), "mask", rel(b["fusion"][1], a["fusion"][1]))
for k, v in a["fusion"][2].items():
    if v is not None: print("fusion", k, rel(b["fusion"][2][k], v))
