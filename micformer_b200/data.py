"""The data side of the hot path: the batch contract of ``dataset/MMWHS.py:308-405`` and the MONAI
transform chain of ``train_mmwhs_noPad.py:116-130``, restated so that they run on whatever device the tensors live on
(on the GPU they are a handful of elementwise torch kernels in front of the step -- plumbing, not the product).

Contract of one sample (``MMWHS.__getitem__``): ``image`` (2, 128, 128, 128) float16 -- channel 0 CT, channel 1 MR, each
min-max or z-score normalised and trilinearly resampled to 128^3; ``label`` (8, 128, 128, 128) bool -- the first 8 planes
of [CT one-hot | MR one-hot] (background + 7 heart structures); plus ``patient_id``, ``seg_path``, ``crop_indexes``,
``et_present``, ``supervised``.  The training loop then does ``.float()`` on both and ``.cuda()`` (``:177-181``).

Transforms (MONAI is absent from this image and the reference pins no version, so these are "parity unpinned"
restatements of the documented semantics, tested against an independent numpy restatement):
  RandFlipd(prob 0.5) on each spatial axis of image AND label; NormalizeIntensityd(nonzero=True, channel_wise=True):
  per channel, over the non-zero voxels only, (x - mean) / std with the population std (1 when it is 0), zeros stay zero;
  RandScaleIntensityd(factors 0.1, prob 1): x *= 1 + U(-0.1, 0.1); RandShiftIntensityd(offsets 0.1, prob 1): x += U(-0.1, 0.1).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def normalize_intensity_nonzero_channelwise(image: Tensor) -> Tensor:
    """NormalizeIntensity(nonzero=True, channel_wise=True) on a (C, D, H, W) or (B, C, D, H, W) tensor (per sample and
    channel).  Returns a new float32 tensor."""
    x = image.float()
    lead = x.shape[:-3]
    flat = x.reshape(-1, x.shape[-3] * x.shape[-2] * x.shape[-1])
    mask = flat != 0
    cnt = mask.sum(1, keepdim=True)
    safe = cnt.clamp(min=1).to(flat.dtype)
    mean = (flat * mask).sum(1, keepdim=True) / safe
    var = (((flat - mean) * mask) ** 2).sum(1, keepdim=True) / safe
    std = var.sqrt()
    std = torch.where(std == 0, torch.ones_like(std), std)
    out = torch.where(mask, (flat - mean) / std, flat)
    return out.reshape(*lead, *x.shape[-3:])


def train_transform(sample: Dict[str, Tensor], generator: Optional[torch.Generator] = None) -> Dict[str, Tensor]:
    """The training chain of ``train_mmwhs_noPad.py:116-125`` on one sample dict (``image`` (C,D,H,W), ``label`` (K,D,H,W)).
    Random draws come from ``generator`` (CPU generator; the draws are scalars, the work happens on the tensors' device)."""
    img, lab = sample["image"], sample["label"]
    for axis in (0, 1, 2):                                             # spatial_axis k == tensor dim k + 1
        if float(torch.rand((), generator=generator)) < 0.5:
            img = torch.flip(img, dims=(axis + 1,))
            lab = torch.flip(lab, dims=(axis + 1,))
    img = normalize_intensity_nonzero_channelwise(img)
    factor = float(torch.rand((), generator=generator)) * 0.2 - 0.1     # U(-0.1, 0.1)
    img = img * (1.0 + factor)
    offset = float(torch.rand((), generator=generator)) * 0.2 - 0.1
    img = img + offset
    out = dict(sample)
    out["image"], out["label"] = img, lab
    return out


def val_transform(sample: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """``:126-130``: intensity normalisation only"""
    out = dict(sample)
    out["image"] = normalize_intensity_nonzero_channelwise(sample["image"])
    return out


class SyntheticMMWHS(torch.utils.data.Dataset):
    """Synthetic stand-in for ``dataset/MMWHS.py`` with the same sample contract (no network / no MM-WHS files here):
    two smooth-ish random modalities with a zero background shell and a random 8-class label volume, deterministic per
    index.  ``transform`` is applied like the reference's (a callable on the sample dict)."""

    def __init__(self, n: int = 16, size: int = 128, num_classes: int = 8, seed: int = 0, transform=None):
        self.n, self.size, self.num_classes, self.seed, self.transform = int(n), int(size), int(num_classes), int(seed), transform

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, idx: int) -> Dict[str, object]:
        if not 0 <= idx < self.n:
            raise IndexError(idx)
        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        S = self.size
        img = torch.rand(2, S, S, S, generator=g)                       # min-max normalised intensities in [0, 1)
        body = torch.zeros(S, S, S, dtype=torch.bool)
        m = max(1, S // 16)
        body[m:S - m, m:S - m, m:S - m] = True                          # zero background shell, as after cropping
        img = (img * body).to(torch.float16)
        cls = torch.randint(0, self.num_classes, (S, S, S), generator=g)
        cls = torch.where(body, cls, torch.zeros_like(cls))
        label = torch.nn.functional.one_hot(cls, self.num_classes).permute(3, 0, 1, 2).contiguous().bool()
        sample = dict(patient_id=f"synthetic_{idx:04d}", image=img, label=label, seg_path="", et_present=0, supervised=True,
                      crop_indexes=((m, S - m), (m, S - m), (m, S - m)))
        return self.transform(sample) if self.transform is not None else sample
