"""Validation-time callers of the hot path: what ``MicFormer/utils.py:222-240`` (``inference``) and
``train_mmwhs_noPad.py:283-306,392-407`` (``meandice``) do around ``model(x)``.

* ``sliding_window_inference`` restates the algorithm of ``monai.inferers.sliding_window_inference`` for the arguments the
  reference passes (``roi_size=(128,128,128)``, ``sw_batch_size=1``, ``overlap=0.5``, default ``mode="constant"``):
  windows start every ``int(roi * (1 - overlap))`` voxels per axis, the last one is clamped to the volume end, predictions
  are averaged with a constant importance map.  MONAI is not installed in this image and the reference pins no version,
  so this piece is "parity unpinned": the tests check its defining properties (single window == predictor, partition of
  unity, window lattice), not MONAI outputs.
* ``meandice`` is the reference's own integer metric (mean Dice over classes 1..C-1 of argmax masks, smooth 1e-6), computed
  from one confusion histogram instead of a Python loop over classes.
* ``inference`` is the reference's wrapper: evaluation mode semantics are the caller's; fp16 inputs (the script validates
  under ``torch.cuda.amp.autocast``) are accepted because ``Head.forward`` up-casts them.

torch is plumbing here (padding, slicing, accumulation); the predictor is the module whose forward runs the sm_100a kernels.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _scan_starts(image: int, roi: int, overlap: float) -> List[int]:
    """window start offsets along one axis: step int(roi*(1-overlap)) (at least 1), last window clamped to the end"""
    if image <= roi:
        return [0]
    step = max(int(roi * (1.0 - overlap)), 1)
    n = -(-(image - roi) // step) + 1                     # ceil((image - roi) / step) + 1
    return [min(i * step, image - roi) for i in range(n)]


def window_lattice(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[Tuple[int, ...]]:
    """start corners of all windows, last axis fastest (the order the windows are evaluated in)"""
    axes = [_scan_starts(i, r, overlap) for i, r in zip(image_size, roi_size)]
    out: List[Tuple[int, ...]] = [()]
    for a in axes:
        out = [o + (s,) for o in out for s in a]
    return out


@torch.no_grad()
def sliding_window_inference(inputs: Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[Tensor], Tensor], overlap: float = 0.5) -> Tensor:
    """inputs (B, C, D, H, W) -> (B, C_out, D, H, W): the predictor is applied to roi-sized windows and the overlapping
    predictions are averaged.  Volumes smaller than the roi are zero-padded symmetrically (MONAI's default) and cropped back."""
    if inputs.ndim != 5:
        raise ValueError("sliding_window_inference: expected a (B, C, D, H, W) tensor")
    if not 0.0 <= overlap < 1.0:
        raise ValueError("overlap must be in [0, 1)")
    roi = tuple(int(r) for r in roi_size)
    orig = tuple(inputs.shape[2:])
    pads = []
    for k in (2, 1, 0):                                   # F.pad takes the last axis first
        diff = max(roi[k] - orig[k], 0)
        pads += [diff // 2, diff - diff // 2]
    x = F.pad(inputs, pads) if any(pads) else inputs
    size = tuple(x.shape[2:])
    B = x.shape[0]
    corners = window_lattice(size, roi, overlap)
    jobs = [(b, c) for b in range(B) for c in corners]    # batch-major, window-minor (sw_batch_size slices this list)
    out = count = None
    for i in range(0, len(jobs), max(1, int(sw_batch_size))):
        chunk = jobs[i:i + max(1, int(sw_batch_size))]
        win = torch.cat([x[b:b + 1, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] for b, (z, y, w) in chunk], 0)
        pred = predictor(win)
        if out is None:
            out = torch.zeros((B, pred.shape[1], *size), dtype=pred.dtype, device=pred.device)
            count = torch.zeros((B, 1, *size), dtype=pred.dtype, device=pred.device)
        for k, (b, (z, y, w)) in enumerate(chunk):
            out[b, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] += pred[k]
            count[b, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] += 1
    out = out / count
    if any(pads):
        z0, y0, w0 = pads[4], pads[2], pads[0]
        out = out[:, :, z0:z0 + orig[0], y0:y0 + orig[1], w0:w0 + orig[2]]
    return out


def inference(input: Tensor, model: Callable[[Tensor], Tensor]) -> Tensor:
    """``utils.py:226-240``: full-volume prediction with 128^3 windows at 50 % overlap, one window per call"""
    return sliding_window_inference(input, (128, 128, 128), 1, model, overlap=0.5)


def meandice(pred: Tensor, label: Tensor, num_class: int) -> Tensor:
    """``train_mmwhs_noPad.py:392-407``: mean over classes 1..num_class-1 of (2|P∩L| + 1e-6) / (|P| + |L| + 1e-6), with
    the sums taken over the whole batch; ``pred`` / ``label`` are integer class maps of the same shape.  The background
    class 0 is excluded, and a class absent from both maps scores 1 (the smooth term), as in the reference."""
    if pred.shape != label.shape:
        raise ValueError(f"meandice: shapes differ {tuple(pred.shape)} vs {tuple(label.shape)}")
    p = pred.reshape(-1).long()
    l = label.reshape(-1).long()
    n = int(num_class)
    inside = (p >= 0) & (p < n) & (l >= 0) & (l < n)
    conf = torch.bincount((l[inside] * n + p[inside]), minlength=n * n).view(n, n).double()   # [label, pred]
    # values outside [0, n) never equal a class index: they count for neither |P| nor |L| of any class
    inter = conf.diagonal()[1:]
    psum = torch.bincount(p[(p >= 0) & (p < n)], minlength=n).double()[1:]
    lsum = torch.bincount(l[(l >= 0) & (l < n)], minlength=n).double()[1:]
    dice = (2.0 * inter + 1e-6) / (psum + lsum + 1e-6)
    return dice.sum() / (n - 1)


def evaluate(logits: Tensor, labels_onehot: Tensor, num_class: int = 8) -> Tensor:
    """the metric line of the validation loop (``:300``): argmax(softmax(logits)) vs argmax(labels) -> meandice"""
    return meandice(torch.argmax(torch.softmax(logits.float(), dim=1), dim=1), torch.argmax(labels_onehot.int(), dim=1), num_class)
