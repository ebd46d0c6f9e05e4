"""B200-native drop-in for the reference's ``dataset`` module (code/dataset.py).

``DWINormalize`` / ``DCENormalize`` keep the reference call signature (a callable on one
[C,H,W] image, code/dataset.py:9-53) and add a batched entry ``batch(x[B,C,H,W])`` - one
kernel launch for a whole batch, which is how the B200 pipeline uses them (normalisation
happens on the device after collation instead of per sample in DataLoader workers).
The arithmetic runs in csrc/normalize.cu; there is no CPU implementation here: CPU
tensors are staged through the GPU and returned on their original device.
The dataset / fold-splitting classes are host-side indexing and keep the reference behaviour.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

import b200_native as nat

__all__ = ["DWINormalize", "DCENormalize", "Resize", "BatchAugment", "SingleInputDataset", "LoadedFusionDataset", "data_segmentation",
           "data_segmentation_mask"]


def _to_device(img):
    if img.is_cuda:
        return img, None
    if not torch.cuda.is_available():
        raise nat.B200NativeError("normalisers run on the GPU only (no CPU path); no CUDA device is visible")
    return img.cuda(non_blocking=True), img.device


class DWINormalize(object):
    """Per-channel z-score -> clip -> [0,1]; the last channel is zeroed when adc=True (dataset.py:9-41)."""

    def __init__(self, clip_z=(-3, 3), adc=True):
        self.z_lo, self.z_hi = clip_z
        self.adc = adc

    def batch(self, x, plane_mean=None):
        """x [B,C,H,W] fp32 CUDA -> normalised fp32 [B,C,H,W]; optionally fills plane_mean [B*C]."""
        if x.dim() != 4:
            raise ValueError("expected [B,C,H,W]")
        x = x.contiguous().float()
        out = torch.empty_like(x)
        B, C, H, W = x.shape
        nat.dwi_normalize(x, out, C, H * W, self.adc, self.z_lo, self.z_hi, plane_mean)
        return out

    def fused_params(self, x):
        """Statistics only, for the encoders' fused first layer (b200_stem_ex applies the map while loading the raw
        planes, so the normalised tensor is never written): -> (("dwi", stats [B*C,4], z_lo, z_hi), plane_mean [B*C] of
        the NORMALISED values).  Register-resident planes only (<= 8 192 samples, multiple of 4)."""
        x = x.contiguous().float()
        B, C, H, W = x.shape
        stats = torch.empty((B * C, 4), dtype=torch.float32, device=x.device)
        pm = torch.empty(B * C, dtype=torch.float32, device=x.device)
        nat._call("b200_dwi_normalize_ex", None, nat._ptr(x), None, B * C, C, H * W, 1 if self.adc else 0, float(self.z_lo),
                  float(self.z_hi), nat._ptr(pm), nat._ptr(stats), nat._stream())
        return ("dwi", stats, float(self.z_lo), float(self.z_hi)), pm

    def __call__(self, img):
        dev_img, home = _to_device(img)
        out = self.batch(dev_img.unsqueeze(0))[0]
        return out if home is None else out.to(home)


class Resize(object):
    """torchvision `transforms.Resize(input_size)` as the reference applies it BEFORE the normaliser
    (code/prepare_single_model.py:112-120): bilinear, align_corners=False, antialias on.  For an upsample
    (64 -> 224, config C4) the antialias filter is the plain bilinear kernel (`b200_resize_bilinear_c1`); where a
    side shrinks (ROIs larger than `input_size`) the filter widens to ATen's triangle of support in / out
    (`b200_resize_aa_c1`).  An int `size` is a square target here, as every ROI of the path is square."""

    def __init__(self, size):
        self.size = (size, size) if isinstance(size, int) else tuple(size)

    def batch(self, x):
        """x [B,C,H,W] fp32 CUDA -> [B,C,size,size] fp32 (one launch, every plane independent)."""
        if x.dim() != 4:
            raise ValueError("expected [B,C,H,W]")
        B, C, H, W = x.shape
        S0, S1 = self.size
        if (S0, S1) == (H, W):
            return x
        x = x.contiguous().float()
        out = torch.empty((B, C, S0, S1), dtype=torch.float32, device=x.device)
        kernel = nat.resize_aa_c1 if (S0 < H or S1 < W) else nat.resize_bilinear_c1
        kernel(x.view(B * C, H, W), out.view(B * C, S0, S1))
        return out

    def __call__(self, img):
        dev_img, home = _to_device(img)
        out = self.batch(dev_img.unsqueeze(0))[0]
        return out if home is None else out.to(home)


class BatchAugment(object):
    """The reference's training augmentation (code/prepare_single_model.py:107-113) -
    `transforms.RandomAffine(degrees=90, translate=(0.1, 0.1), shear=(0.1, 0.1))`, `RandomHorizontalFlip()`,
    `RandomVerticalFlip()` - for a whole batch in one kernel pass on the device.

    Parameters are drawn on the host exactly as torchvision draws them, sample by sample and in the Compose order
    (RandomAffine.get_params: angle, tx, ty, [scale], shear_x[, shear_y]; then one `torch.rand(1)` per flip), from
    torch's global CPU generator - so under the same seed a batch gets the parameters the reference's per-sample
    pipeline would give the same images.  `inverse_matrix` restates
    torchvision.transforms.functional._get_inverse_affine_matrix for the tensor backend (centre at the image centre).
    Nearest-neighbour interpolation and zero fill, torchvision's defaults."""

    def __init__(self, degrees=90, translate=(0.1, 0.1), scale=None, shear=(0.1, 0.1), hflip_p=0.5, vflip_p=0.5,
                 fill=0.0):
        self.degrees = (-float(degrees), float(degrees)) if isinstance(degrees, (int, float)) else tuple(degrees)
        self.translate, self.scale = translate, scale
        if isinstance(shear, (int, float)):
            shear = (-float(shear), float(shear))
        self.shear = None if shear is None else tuple(float(s) for s in shear)
        if self.shear is not None and len(self.shear) not in (2, 4):
            raise ValueError("shear must hold 2 or 4 values")
        self.hflip_p, self.vflip_p, self.fill = hflip_p, vflip_p, float(fill)

    def sample_params(self, n, height, width):
        """-> list of n (angle, (tx, ty), scale, (shear_x, shear_y), hflip, vflip), torchvision's draw order."""
        out = []
        for _ in range(n):
            angle = float(torch.empty(1).uniform_(self.degrees[0], self.degrees[1]).item())
            tx = ty = 0
            if self.translate is not None:
                max_dx, max_dy = float(self.translate[0] * width), float(self.translate[1] * height)
                tx = int(round(torch.empty(1).uniform_(-max_dx, max_dx).item()))
                ty = int(round(torch.empty(1).uniform_(-max_dy, max_dy).item()))
            sc = 1.0
            if self.scale is not None:
                sc = float(torch.empty(1).uniform_(self.scale[0], self.scale[1]).item())
            shx = shy = 0.0
            if self.shear is not None:
                shx = float(torch.empty(1).uniform_(self.shear[0], self.shear[1]).item())
                if len(self.shear) == 4:
                    shy = float(torch.empty(1).uniform_(self.shear[2], self.shear[3]).item())
            hf = bool(torch.rand(1) < self.hflip_p) if self.hflip_p is not None else False
            vf = bool(torch.rand(1) < self.vflip_p) if self.vflip_p is not None else False
            out.append((angle, (tx, ty), sc, (shx, shy), hf, vf))
        return out

    @staticmethod
    def inverse_matrix(angle, translate, scale, shear):
        """Inverse affine matrix (output -> input pixel, centred coordinates): rotation-scale-shear inverse times the
        inverse translation, float64 arithmetic like torchvision's."""
        import math

        rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
        tx, ty = float(translate[0]), float(translate[1])
        a = math.cos(rot - sy) / math.cos(sy)
        b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
        c = math.sin(rot - sy) / math.cos(sy)
        d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
        m = [d / scale, -b / scale, 0.0, -c / scale, a / scale, 0.0]
        m[2] += m[0] * (-tx) + m[1] * (-ty)
        m[5] += m[3] * (-tx) + m[4] * (-ty)
        return m

    def batch(self, x, params=None):
        """x [B,C,H,W] fp32 CUDA -> augmented copy.  `params` (from sample_params) may be passed to reuse a draw."""
        if x.dim() != 4:
            raise ValueError("expected [B,C,H,W]")
        if not x.is_cuda:
            raise nat.B200NativeError("BatchAugment.batch needs a CUDA tensor (no CPU path)")
        B, C, H, W = x.shape
        if params is None:
            params = self.sample_params(B, H, W)
        theta = torch.tensor([self.inverse_matrix(p[0], p[1], p[2], p[3]) for p in params], dtype=torch.float32)
        flips = torch.tensor([int(p[4]) | (int(p[5]) << 1) for p in params], dtype=torch.int32)
        x = x.contiguous().float()
        out = torch.empty_like(x)
        nat.augment(x, theta.to(x.device), flips.to(x.device), out, self.fill)
        return out

    def __call__(self, img):
        dev_img, home = _to_device(img)
        out = self.batch(dev_img.unsqueeze(0))[0]
        return out if home is None else out.to(home)


class DCENormalize(object):
    """Nyul standardisation of one image through the fitted standardiser (dataset.py:46-53)."""

    def __init__(self, nyul_standardizer):
        self.nyul = nyul_standardizer

    def batch(self, x, plane_mean=None):
        return self.nyul.transform_batch(x, plane_mean=plane_mean)

    def fused_params(self, x):
        """Per-plane composed tables only (see DWINormalize.fused_params): -> (("nyul", tables [B*C,56] fp64, L),
        plane_mean [B*C] of the standardised values)."""
        return self.nyul.tables_batch(x)

    def __call__(self, img):
        norm = self.nyul.transform(img)
        return norm.to(img.device)


class SingleInputDataset(torch.utils.data.Dataset):
    """(img[, mask][, label]) samples with optional transforms and ADC channel (dataset.py:56-98)."""

    def __init__(self, imgs, masks=None, labels=None, transforms=None, modality="dwi", nyul_standardizer=None,
                 adc_min=None, adc_map=None):
        self.imgs, self.masks, self.labels = imgs, masks, labels
        self.transforms, self.modality, self.adc_map = transforms, modality, adc_map

    def __len__(self):
        return len(self.imgs)

    def __getitem__(self, index):
        img = self.imgs[index].clone()
        label = self.labels[index] if self.labels is not None else None
        mask = self.masks[index] if self.masks is not None else None
        if self.transforms:
            img = self.transforms(img)
        if self.adc_map is not None:  # the ADC channel is resized to the (transformed) image (reference :79-88)
            adc = self.adc_map
            if tuple(adc.shape[-2:]) != tuple(img.shape[-2:]):
                dev_adc, home = _to_device(adc)
                out = torch.empty((adc.shape[0], *img.shape[-2:]), dtype=torch.float32, device=dev_adc.device)
                nat.resize_bilinear_c1(dev_adc.contiguous().float(), out)
                adc = out if home is None else out.to(home)
            img = torch.cat([img, adc.to(img.device)], dim=0)
        items = [img.float()]
        if mask is not None:
            items.append(mask.float())
        if label is not None:
            items.append(label)
        return items[0] if len(items) == 1 else tuple(items)


class LoadedFusionDataset(torch.utils.data.Dataset):
    """Pre-processed (dwi, dce[, mask][, label]) tuples (dataset.py:100-140)."""

    def __init__(self, dwi, dce, masks=None, labels=None):
        self.dwi, self.dce, self.masks, self.labels = dwi, dce, masks, labels
        self.length = len(dwi)
        assert len(dwi) == len(dce), "DWI and DCE must have same length"
        if masks is not None:
            assert len(masks) == len(dwi), "Masks must match DWI length"
        if labels is not None:
            assert len(labels) == len(dwi), "Labels must match DWI length"

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        x1, x2 = self.dwi[index].clone().float(), self.dce[index].clone().float()
        m = self.masks[index] if self.masks is not None else None
        y = self.labels[index] if self.labels is not None else None
        if m is not None and y is not None:
            return x1, x2, m.float(), y
        if y is not None:
            return x1, x2, y
        if m is not None:
            return x1, x2, m.float()
        return x1, x2


def _fold_indices(labels, segnum, classnum, fold):
    """Index form of the reference's stratified K-fold (dataset.py:142-176, :178-235): seed numpy once
    with 42, permute every class's indices in class order, cut each class into `segnum` slices of
    floor(n/segnum) (the last slice takes the remainder), segment i = concat over classes of slice i;
    segment `fold` is validation, the others (in order) are training."""
    np.random.seed(42)
    per_class = []
    for c in range(classnum):
        idx = torch.where(labels == c)[0]
        per_class.append(idx[np.random.permutation(idx.size(0))].tolist())
    segments = []
    for i in range(segnum):
        seg = []
        for idx in per_class:
            step = len(idx) // segnum
            seg += idx[i * step:(i + 1) * step] if i != segnum - 1 else idx[(segnum - 1) * step:]
        segments.append(seg)
    train = [j for i, seg in enumerate(segments) if i != fold for j in seg]
    return train, segments[fold]


def data_segmentation(imgs, labels, segnum, classnum, fold):
    tr, va = _fold_indices(labels, segnum, classnum, fold)
    return [imgs[tr], imgs[va]], [labels[tr].float(), labels[va].float()]


def data_segmentation_mask(imgs, masks, labels, segnum, classnum, fold):
    tr, va = _fold_indices(labels, segnum, classnum, fold)
    return [imgs[tr], imgs[va]], [masks[tr], masks[va]], [labels[tr], labels[va]]
