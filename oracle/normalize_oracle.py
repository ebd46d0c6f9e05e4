"""Oracle for the input normalisers (test infrastructure only - see oracle/__init__.py).

Restates, operation by operation and dtype by dtype:
  * DWINormalize.__call__            /root/reference/code/dataset.py:14-41
  * NyulStandardizer.fit / transform /root/reference/code/preprocess_helpers.py:65-120
  * DCE pre-scaling by the case max  /root/reference/code/prepare_single_model.py:337-343
  * compute_adc_map                  /root/reference/code/preprocess_helpers.py:133-167
"""
from __future__ import annotations

import numpy as np
import torch

DEFAULT_LANDMARKS = (1, 10, 25, 30, 40, 50, 60, 75, 80, 90, 99)  # preprocess_helpers.py:53


def dwi_normalize(img: torch.Tensor, clip_z=(-3.0, 3.0), adc: bool = True) -> torch.Tensor:
    """dataset.py:14-41.  img [C,H,W] fp32; channel C-1 stays zero when adc (dataset.py:17-23)."""
    z_lo, z_hi = clip_z
    n_ch = img.shape[0]
    out = torch.zeros_like(img)
    for ch in range(n_ch - 1 if adc else n_ch):
        x = img[ch]
        mean = x.mean()                       # dataset.py:28
        std = x.std().clamp(min=1e-6)         # dataset.py:29 (unbiased)
        x = (x - mean) / std                  # dataset.py:30
        x = torch.clamp(x, z_lo, z_hi)        # dataset.py:33
        out[ch] = (x - z_lo) / (z_hi - z_lo)  # dataset.py:36
    return out


def dwi_normalize_batch(x: torch.Tensor, clip_z=(-3.0, 3.0), adc: bool = True) -> torch.Tensor:
    return torch.stack([dwi_normalize(c, clip_z, adc) for c in x])


def percentile_linear(sorted_x: np.ndarray, landmarks) -> np.ndarray:
    """np.percentile(x, landmarks) ("linear" rule) written out: preprocess_helpers.py:62-63, :100.

    numpy computes the neighbour difference in the array dtype (float32) and the lerp in
    float64, switching formula at gamma >= 0.5 (numpy/lib/_function_base_impl.py:_lerp).
    """
    n = sorted_x.shape[0]
    q = np.true_divide(np.asarray(landmarks, dtype=np.float64), 100.0)
    virt = (n - 1) * q
    prev = np.floor(virt).astype(np.intp)
    nxt = np.minimum(prev + 1, n - 1)
    gamma = virt - prev
    a, b = sorted_x[prev], sorted_x[nxt]
    diff = b - a  # float32
    lo = a.astype(np.float64) + diff.astype(np.float64) * gamma
    hi = b.astype(np.float64) - diff.astype(np.float64) * (1.0 - gamma)
    return np.where(gamma >= 0.5, hi, lo)


def percentile_indices(n: int, landmarks):
    """(prev_index int32 [L], gamma float64 [L]) of the linear percentile rule for n samples."""
    q = np.true_divide(np.asarray(landmarks, dtype=np.float64), 100.0)
    virt = (n - 1) * q
    prev = np.floor(virt)
    return prev.astype(np.int32), (virt - prev).astype(np.float64)


def nyul_fit(images, num_channels: int = 6, landmarks=DEFAULT_LANDMARKS) -> np.ndarray:
    """preprocess_helpers.py:65-83: mean over images of each channel's landmark vector -> [C,L] float64."""
    acc = [[] for _ in range(num_channels)]
    for img in images:
        arr = img.cpu().numpy() if torch.is_tensor(img) else np.asarray(img)
        for c in range(num_channels):
            acc[c].append(percentile_linear(np.sort(arr[c].reshape(-1)), landmarks))
    return np.stack([np.mean(acc[c], axis=0) for c in range(num_channels)])


def nyul_transform(img, channel_landmarks: np.ndarray, landmarks=DEFAULT_LANDMARKS, target_range=(0.0, 1.0)):
    """preprocess_helpers.py:85-120.  img [C,H,W] fp32 (tensor or array) -> same type, fp32."""
    is_tensor = torch.is_tensor(img)
    arr = img.cpu().numpy() if is_tensor else np.asarray(img)
    scale = np.linspace(target_range[0], target_range[1], len(landmarks))  # :60
    out = np.zeros_like(arr, dtype=np.float32)
    for c in range(channel_landmarks.shape[0]):
        flat = arr[c].reshape(-1)
        orig = percentile_linear(np.sort(flat), landmarks)  # :100
        avg = channel_landmarks[c]                          # :103
        mid = np.interp(flat, orig, avg)                    # :105
        mid = np.interp(mid, avg, scale)                    # :108
        out[c] = mid.reshape(arr[c].shape)                  # :111
    return torch.tensor(out, dtype=torch.float32) if is_tensor else out


def nyul_transform_batch(x: torch.Tensor, channel_landmarks: np.ndarray, **kw) -> torch.Tensor:
    return torch.stack([nyul_transform(c, channel_landmarks, **kw) for c in x])


def dce_prescale(x: torch.Tensor) -> torch.Tensor:
    """prepare_single_model.py:337-343: divide every case by its max over channels and pixels."""
    return x / x.amax(dim=(1, 2, 3), keepdim=True)


def compute_adc_map(dwi: torch.Tensor, bvals, eps: float = 1e-6) -> torch.Tensor:
    """preprocess_helpers.py:133-167: -slope of the least-squares line of log S against b."""
    n_ch = dwi.shape[0]
    b = torch.tensor(bvals, dtype=torch.float32).view(n_ch, 1, 1)
    log_s = torch.log(torch.clamp(dwi, min=eps))
    cov = ((b - b.mean()) * (log_s - log_s.mean(dim=0))).sum(dim=0)
    var = ((b - b.mean()) ** 2).sum()
    return (-(cov / (var + eps))).unsqueeze(0)


def resize(x: torch.Tensor, size: int) -> torch.Tensor:
    """torchvision `transforms.Resize(size)` on a float tensor [..., H, W] as the reference applies it ahead of
    the normaliser (code/prepare_single_model.py:112, :116, :120): bilinear, align_corners=False, antialias=True
    (torchvision's default for tensors since 0.17).  tests/test_oracle_golden.py checks it against
    torchvision.transforms.Resize itself."""
    lead = x.shape[:-2]
    y = torch.nn.functional.interpolate(x.reshape(-1, 1, *x.shape[-2:]).float(), size=(size, size), mode="bilinear",
                                        align_corners=False, antialias=True)
    return y.reshape(*lead, size, size)
