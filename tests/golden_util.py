"""Helpers to compare tensors with the committed reference fingerprints (tests/golden/*.npz)."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROBE_STRIDE = 997
PROBE_N = 256


def load(name):
    return np.load(os.path.join(GOLD, name))


def load_shapes(tag):
    with open(os.path.join(GOLD, f"state_shapes_{tag}.json")) as f:
        raw = json.load(f)
    return {m: {k: tuple(v) for k, v in d.items()} for m, d in raw.items()}


def to_nchw(t):
    """Product maps are channels-last bf16 views with NCHW shape; fingerprints index NCHW-contiguous order."""
    return t.detach().float().cpu().contiguous()


def check(gold, key, t, rtol, atol_frac=None, what=""):
    """Compare tensor `t` with fingerprint `key`.  Error metric: max|d| / max|ref| over the probe
    (and over the full tensor when the fixture stores it)."""
    t = to_nchw(t)
    meta = gold[key + "/meta"]
    shape = tuple(int(v) for v in meta[:-1])
    assert tuple(t.shape) == shape, f"{key}: shape {tuple(t.shape)} != {shape}"
    if key + "/full" in gold.files:
        # the whole tensor is stored: the metric is exact.  (The strided probe of a tensor smaller than the stride
        # is its first element alone, and max|d| / |ref[0]| is not the metric - a 2 x 1 x 14 x 14 map whose first
        # element is near zero reads as 20 % off while every element is within 2 % of the map's range.)
        ref_f = torch.from_numpy(gold[key + "/full"]).double()
        err = (t.double() - ref_f).abs().max().item() / max(ref_f.abs().max().item(), 1e-12)
    else:
        ref_s = torch.from_numpy(gold[key + "/samples"]).double()
        got_s = t.double().reshape(-1)[::PROBE_STRIDE][:PROBE_N]
        err = (got_s - ref_s).abs().max().item() / max(ref_s.abs().max().item(), 1e-12)
    assert err <= rtol, f"{what}{key}: max|d|/max|ref| = {err:.3e} > {rtol:.1e}"
    ref_sum, ref_abs = gold[key + "/sums"]
    assert abs(t.double().sum().item() - ref_sum) <= 4 * rtol * max(ref_abs, 1e-12) + 1e-9, f"{key}: sum mismatch"
    return err


def walk(prefix, obj):
    """Yield (key, tensor) with the naming oracle/make_golden.py uses."""
    if obj is None:
        return
    if torch.is_tensor(obj):
        yield prefix, obj
    elif isinstance(obj, (list, tuple)):
        for i, o in enumerate(obj):
            yield from walk(f"{prefix}.{i}", o)
    elif isinstance(obj, dict):
        for k, o in obj.items():
            yield from walk(f"{prefix}.{k}", o)
