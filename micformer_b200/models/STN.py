"""Drop-in for the reference's ``MicFormer/models/STN.py``.

``SpatialTransformer.forward(src, flow)`` (STN.py:9-32): ``src`` (B,C,D,H,W), ``flow`` (B,3,D,H,W) voxel
displacements; returns ``grid_sample(src, 2*((idx+flow)/(S-1) - .5)[..., (2,1,0)], bilinear, zeros,
align_corners=False)``.  Inside the cross block the resampling is fused with its neighbours; this module is the
stand-alone entry (channels-first in/out like the reference, permuted around the channels-last kernel).
"""
import torch
import torch.nn as nn

from .. import ops


class SpatialTransformer(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, src, flow, mode='bilinear'):
        if mode != 'bilinear':
            raise NotImplementedError("micformer_b200: only mode='bilinear' (the reference's only call) is built")
        if src.dim() != 5 or flow.dim() != 5 or flow.shape[1] != 3:
            raise NotImplementedError("micformer_b200: SpatialTransformer is built for 3-D volumes only")
        C = src.shape[1]
        pad = (-C) % 4
        s = src.permute(0, 2, 3, 4, 1)
        if pad:
            s = torch.nn.functional.pad(s, (0, pad))
        out = ops.DeformSampleFn.apply(s.contiguous(), flow.permute(0, 2, 3, 4, 1).contiguous())
        return out[..., :C].permute(0, 4, 1, 2, 3)


class Re_SpatialTransformer(nn.Module):
    """STN.py:35-43 (dead code in the reference; kept for import parity)."""

    def __init__(self):
        super().__init__()
        self.stn = SpatialTransformer()

    def forward(self, src, flow, mode='bilinear'):
        flow = -1 * self.stn(flow, flow, mode='bilinear')
        return self.stn(src, flow, mode)
