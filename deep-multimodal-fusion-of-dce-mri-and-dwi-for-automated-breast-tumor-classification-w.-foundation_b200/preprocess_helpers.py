"""B200-native drop-in for the reference's ``preprocess_helpers`` (code/preprocess_helpers.py).

``NyulStandardizer`` keeps the reference interface (``fit``, ``transform``, ``save``, ``load``,
``fitted``, ``channel_landmarks``, code/preprocess_helpers.py:52-130).  ``transform`` and the
batched ``transform_batch`` run csrc/normalize.cu's nyul_transform kernel (in-SM sort for the
11 order statistics, float64 piece-wise linear maps with numpy.interp's branch structure).
``fit`` is a one-off host step over the training set and stays in numpy.
"""
from __future__ import annotations

import numpy as np
import torch

import b200_native as nat

__all__ = ["NyulStandardizer", "preprocess_dce", "prescale_dce", "prep_data_by_mod", "zero_to_one_adc", "normalize_adc", "preprocess_adc",
           "compute_adc_map"]


class NyulStandardizer:
    def __init__(self, landmarks=[1, 10, 25, 30, 40, 50, 60, 75, 80, 90, 99], target_range=(0, 1)):
        self.landmarks = landmarks
        self.fitted = False
        self.channel_landmarks = None
        self.standard_scale = np.linspace(target_range[0], target_range[1], len(landmarks))
        self._dev_cache = {}

    def _percentiles(self, img_np):
        return np.percentile(img_np.flatten(), self.landmarks)

    def fit(self, images, num_channels=6):
        """Average, over the training images, of every channel's landmark vector (:65-83)."""
        per_channel = [[] for _ in range(num_channels)]
        for img in images:
            arr = img.detach().cpu().numpy() if torch.is_tensor(img) else np.asarray(img)
            for c in range(num_channels):
                per_channel[c].append(self._percentiles(arr[c]))
        self.channel_landmarks = {c: np.mean(per_channel[c], axis=0) for c in range(num_channels)}
        self.fitted = True
        self._dev_cache = {}

    def _device_tables(self, dev, n, num_channels):
        key = (str(dev), n, num_channels)
        if key not in self._dev_cache:
            q = np.true_divide(np.asarray(self.landmarks, dtype=np.float64), 100.0)
            virt = (n - 1) * q  # numpy's "linear" virtual index
            prev = np.floor(virt)
            lm = np.stack([np.asarray(self.channel_landmarks[c], dtype=np.float64) for c in range(num_channels)])
            self._dev_cache[key] = (
                torch.from_numpy(lm).to(dev), torch.from_numpy(np.asarray(self.standard_scale, np.float64)).to(dev),
                torch.from_numpy(prev.astype(np.int32)).to(dev), torch.from_numpy(virt - prev).to(dev))
        return self._dev_cache[key]

    def transform_batch(self, x, num_channels=None, plane_mean=None, exact=False):
        """x [B,C,H,W] fp32 CUDA -> standardised fp32 [B,C,H,W].  `exact=True` reproduces numpy's fp64 operation
        order bit for bit; the default composes the two interpolations into one table per plane (<= 1 fp32 ulp)."""
        if not self.fitted:
            raise RuntimeError("Call fit() first")
        x = x.contiguous().float()
        B, C, H, W = x.shape
        num_channels = C if num_channels is None else num_channels
        if num_channels != C:
            raise ValueError("transform_batch standardises every channel of the batch")
        avg, scale, prev, gamma = self._device_tables(x.device, H * W, C)
        out = torch.empty_like(x)
        nat.nyul_transform(x, out, C, H * W, avg, scale, prev, gamma, plane_mean, exact=exact)
        return out

    def tables_batch(self, x):
        """The per-plane composed piece-wise linear tables of `transform_batch` WITHOUT applying them (the encoders'
        fused first layer does that while loading the raw planes): -> (("nyul", tables [B*C,56] fp64, L), plane_mean
        [B*C] fp32 of the standardised values)."""
        if not self.fitted:
            raise RuntimeError("Call fit() first")
        x = x.contiguous().float()
        B, C, H, W = x.shape
        avg, scale, prev, gamma = self._device_tables(x.device, H * W, C)
        tables = torch.zeros((B * C, 56), dtype=torch.float64, device=x.device)
        pm = torch.empty(B * C, dtype=torch.float32, device=x.device)
        L = scale.numel()
        nat._call("b200_nyul_transform_ex2", None, nat._ptr(x), None, B * C, C, H * W, L, nat._ptr(avg), nat._ptr(scale),
                  nat._ptr(prev), nat._ptr(gamma), nat._ptr(pm), 0, nat._ptr(tables), nat._stream())
        return ("nyul", tables, L), pm

    def transform(self, img, num_channels=6):
        """One [C,H,W] image (tensor or numpy); channels >= num_channels come back as zeros (:85-120)."""
        if not self.fitted:
            raise RuntimeError("Call fit() first")
        is_tensor = torch.is_tensor(img)
        t = img if is_tensor else torch.from_numpy(np.asarray(img))
        if not torch.cuda.is_available():
            raise nat.B200NativeError("NyulStandardizer.transform runs on the GPU only (no CPU path)")
        dev_t = t.cuda() if not t.is_cuda else t
        out = torch.zeros_like(dev_t, dtype=torch.float32)
        out[:num_channels] = self.transform_batch(dev_t[:num_channels].unsqueeze(0).float())[0]
        out = out.cpu()
        return out if is_tensor else out.numpy()

    def save(self, path):
        np.save(path, {"channel_landmarks": self.channel_landmarks, "fitted": self.fitted})

    def load(self, path):
        data = np.load(path, allow_pickle=True).item()
        self.channel_landmarks = data["channel_landmarks"]
        self.fitted = data["fitted"]
        self._dev_cache = {}


def preprocess_dce(dce_tensor, nyul_model, apply_zscore=False):
    """[C,H,W] -> Nyul-standardised tensor of the input dtype (:5-22)."""
    C = dce_tensor.shape[0]
    out = nyul_model.transform(dce_tensor, num_channels=C)
    if apply_zscore:
        for c in range(C):
            std = out[c].std()
            if std > 1e-8:
                out[c] = (out[c] - out[c].mean()) / std
    return out.clone().to(dce_tensor.dtype) if torch.is_tensor(out) else torch.tensor(out, dtype=dce_tensor.dtype)


def zero_to_one_adc(adc_map, adc_min=None, adc_max=None):
    return ((adc_map - adc_min) / (adc_max - adc_min + 1e-8)).clamp(0, 1)


def normalize_adc(adc_map):
    return adc_map.clamp(0, 3e-3) / 3e-3


def preprocess_adc(adc_map):
    return normalize_adc(torch.log1p(adc_map.clamp(min=0)))


def compute_adc_map(dwi_imgs, bvals, eps=1e-6):
    """-slope of the per-pixel least-squares line of log S over b (:133-167) -> [1,H,W].  Runs on the device
    (b200_adc_map); a CPU tensor is staged through the GPU and returned on the CPU, like the normalisers."""
    home = None if dwi_imgs.is_cuda else dwi_imgs.device
    if home is not None and not torch.cuda.is_available():
        raise nat.B200NativeError("compute_adc_map runs on the GPU only (no CPU path)")
    x = dwi_imgs if dwi_imgs.is_cuda else dwi_imgs.cuda()
    b = torch.as_tensor(bvals, dtype=torch.float32).to(x.device)
    out = nat.adc_map(x.unsqueeze(0), b, eps)[0]
    return out if home is None else out.to(home)


def compute_adc_map_batch(dwi, bvals, eps=1e-6):
    """Batched form: dwi [B,C,H,W] CUDA -> [B,1,H,W] (one launch)."""
    return nat.adc_map(dwi, torch.as_tensor(bvals, dtype=torch.float32).to(dwi.device), eps)


def prescale_dce(imgs):
    """DCE pre-scale of `prep_data_by_mod` (reference prepare_single_model.py:337-343): every case [N,C,H,W] divided by
    its maximum over all channels and pixels - `imgs / imgs_max[:, None, None, None]`.  Runs on the GPU
    (b200_case_max_scale, one CTA per case); a CPU tensor raises (no CPU path)."""
    if not imgs.is_cuda:
        raise nat.B200NativeError("prescale_dce needs a CUDA tensor (no CPU path)")
    x = imgs.contiguous().float()
    out = torch.empty_like(x)
    nat._call("b200_case_max_scale", None, nat._ptr(x), x.shape[0], x[0].numel(), nat._ptr(out), nat._stream())
    return out


def prep_data_by_mod(method, bvals, imgs, test_imgs, parameters):
    """The DCE branch of the reference's `prep_data_by_mod` (prepare_single_model.py:311-343): both splits pre-scaled per
    case; returns (imgs, test_imgs, None) like the reference.  (The DWI branch - ADC maps - is `compute_adc_map` +
    `preprocess_adc` + `zero_to_one_adc` above.)"""
    if method != "dce":
        raise NotImplementedError("prep_data_by_mod: the DCE pre-scale; DWI's ADC maps go through compute_adc_map")
    return prescale_dce(imgs), prescale_dce(test_imgs), None
