"""torch.library layer (b200_ops.py): registration, fake (meta) implementations and graph capture - checked on fake
CUDA tensors, which need no GPU.  The real implementations are covered by tests/test_zz_ops_gpu.py."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode
from torch.fx.experimental.proxy_tensor import make_fx

import b200_native as nat
import b200_ops  # noqa: F401  (registers torch.ops.b200.*)


def test_operators_are_registered_with_shape_correct_fakes():
    ops = torch.ops.b200
    with FakeTensorMode():
        x = torch.empty(4, 16, 64, 64, device="cuda")
        out, pm = ops.dwi_normalize(x, True, -3.0, 3.0)
        assert out.shape == x.shape and pm.shape == (64,) and out.device.type == "cuda"
        c = torch.empty(4, 6, 64, 64, device="cuda")
        tabs = (torch.empty(6, 11, dtype=torch.float64, device="cuda"), torch.empty(11, dtype=torch.float64, device="cuda"),
                torch.empty(22, dtype=torch.int32, device="cuda"), torch.empty(22, dtype=torch.float64, device="cuda"))
        o2, pm2 = ops.nyul_transform(c, *tabs)
        assert o2.shape == c.shape and pm2.shape == (24,)
        assert ops.resize_bilinear(x, 224, 224).shape == (4, 16, 224, 224)
        th, fl = torch.empty(4, 6, device="cuda"), torch.empty(4, dtype=torch.int32, device="cuda")
        assert ops.augment(x, th, fl, 0.0).shape == x.shape
        f = torch.empty(4, 32, 32, 256, device="cuda", dtype=torch.bfloat16)
        w = torch.empty(128, 9 * 256, device="cuda", dtype=torch.bfloat16)
        y = ops.conv_gemm(f, w, None, None, None, 0, 1, 9)
        assert y.shape == (4, 32, 32, 128) and y.dtype == torch.bfloat16
        t = ops.fusion_tokens(y, 4, 4)
        assert t.shape == (4, 16, 128) and t.dtype == torch.float32
        qkv = torch.empty(3 * 197, 3 * 768, device="cuda", dtype=torch.bfloat16)
        o = ops.attention(qkv, 3, 197, 12, 64)
        assert o.shape == (3 * 197, 768) and o.dtype == torch.bfloat16


def test_graph_capture_sees_opaque_b200_nodes():
    def fn(x):
        o, _ = torch.ops.b200.dwi_normalize(x, True, -3.0, 3.0)
        return torch.ops.b200.resize_bilinear(o, 224, 224) * 2

    with FakeTensorMode(allow_non_fake_inputs=True):
        g = make_fx(fn, tracing_mode="fake")(torch.empty(2, 16, 64, 64, device="cuda"))
    targets = [n.target for n in g.graph.nodes if n.op == "call_function"]
    assert torch.ops.b200.dwi_normalize.default in targets and torch.ops.b200.resize_bilinear.default in targets


def test_no_cpu_implementation():
    with pytest.raises(nat.B200NativeError):
        torch.ops.b200.dwi_normalize(torch.zeros(1, 2, 4, 4), True, -3.0, 3.0)
    with pytest.raises(nat.B200NativeError):
        torch.ops.b200.fusion_tokens(torch.zeros(1, 8, 8, 16, dtype=torch.bfloat16), 4, 4)
