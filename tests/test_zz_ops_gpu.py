"""torch.ops.b200.* on the GPU: every operator gives exactly what the package's own call path gives (they launch the
same kernels), eagerly and through torch.compile's graph capture."""
import pytest
import torch

import b200_native as nat
import b200_ops  # noqa: F401
import dataset as ds
import preprocess_helpers as pre
from oracle import params as op

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_operators_equal_the_direct_call_path():
    ops = torch.ops.b200
    dwi, dce, _, _ = op.synthetic_raw(4, seed=8, kind="S")
    dwi, dce = dwi.to(DEV), dce.to(DEV)
    out, pm = ops.dwi_normalize(dwi, True, -3.0, 3.0)
    pm_ref = torch.empty(4 * 16, device=DEV)
    assert torch.equal(out, ds.DWINormalize().batch(dwi, plane_mean=pm_ref)) and torch.equal(pm, pm_ref)
    nyul = pre.NyulStandardizer()
    nyul.fit(list(dce.cpu()), num_channels=6)
    tabs = nyul._device_tables(dce.device, 64 * 64, 6)
    o2, _ = ops.nyul_transform(dce, *tabs)
    assert torch.equal(o2, nyul.transform_batch(dce))
    assert torch.equal(ops.resize_bilinear(dwi, 224, 224), ds.Resize(224).batch(dwi))
    assert torch.equal(ops.resize_bilinear(dwi, 48, 40), ds.Resize((48, 40)).batch(dwi))  # shrinking: antialiased
    aug = ds.BatchAugment()
    torch.manual_seed(2)
    params = aug.sample_params(4, 64, 64)
    theta = torch.tensor([aug.inverse_matrix(*p[:4]) for p in params], dtype=torch.float32, device=DEV)
    flips = torch.tensor([int(p[4]) | (int(p[5]) << 1) for p in params], dtype=torch.int32, device=DEV)
    assert torch.equal(ops.augment(dwi, theta, flips, 0.0), aug.batch(dwi, params=params))
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(2, 32, 32, 128, generator=g, device=DEV).to(torch.bfloat16)
    w = (torch.randn(128, 9 * 128, generator=g, device=DEV) * 0.03).to(torch.bfloat16)
    bias = torch.randn(128, generator=g, device=DEV)
    y = ops.conv_gemm(x, w, None, bias, None, 0, 1, 9)
    assert torch.equal(y, nat.conv_gemm(x, w, taps=9, bias=bias, act=1))
    tok = ops.fusion_tokens(y, 4, 4)
    ref = torch.empty(2, 16, 128, device=DEV)
    nat.fusion_tokens(y, 4, 4, ref)
    assert torch.equal(tok, ref)
    qkv = (torch.randn(2 * 197, 3 * 768, generator=g, device=DEV) * 0.8).to(torch.bfloat16)
    o_ref = torch.empty(2 * 197, 768, device=DEV, dtype=torch.bfloat16)
    nat.attention(qkv, o_ref, 2, 197, 12, 64)
    assert torch.equal(ops.attention(qkv, 2, 197, 12, 64), o_ref)


def test_torch_compile_traces_through_the_operators():
    def fn(x):
        o, pm = torch.ops.b200.dwi_normalize(x, True, -3.0, 3.0)
        return torch.ops.b200.resize_bilinear(o, 128, 128), pm

    x = torch.rand(2, 16, 64, 64, device=DEV) * 500 + 1
    want = fn(x)
    got = torch.compile(fn, backend="eager", fullgraph=True)(x)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_torch_compile_of_the_models_runs_the_kernels():
    """`torch.compile(model, backend='inductor')`, as the reference applies with parameters['compile']
    (run_training.py:90-91, :242-244): the compiled wrappers of both encoders and the fusion head return what the plain
    modules return (their forwards are opaque to Dynamo)."""
    import model_module as mm
    import parameters_default as pd

    torch.manual_seed(0)
    p = pd.default_parameters()
    enc = mm.ModelMaskHeadBackbone("dce", p).to("cuda").eval()
    fus = mm.FusionModel(p).to("cuda").eval()
    x = torch.rand(2, 6, 64, 64, device="cuda")
    with torch.no_grad():
        l0, a0, m0 = enc(x)
        l1, a1, m1 = torch.compile(enc, backend="inductor")(x)
        f0 = fus(a0["raw_feats"], a0["raw_feats"], m0, m0)
        f1 = torch.compile(fus, backend="inductor")(a1["raw_feats"], a1["raw_feats"], m1, m1)
    assert torch.allclose(l0, l1, rtol=1e-3, atol=1e-5) and torch.allclose(m0.float(), m1.float(), rtol=1e-3, atol=1e-4)
    assert torch.allclose(f0[0], f1[0], rtol=1e-3, atol=1e-5)
