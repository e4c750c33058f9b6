"""Drop-in for ``MDiceLoss`` / ``MDiceLoss_Val`` of the reference's ``MicFormer/loss/dice.py:119-230``.

``forward(inputs, target)``: logits (B,C,D,H,W) and one-hot targets of the same shape ->
``(0.7 * sum_c dice_c + 0.3 * sum_c BCE_c) / C`` (``MDiceLoss``, dice.py:158-166) or ``sum_c dice_c / C``
(``MDiceLoss_Val``, dice.py:216-221) with per-channel sums over batch AND space, computed by one fused reduction
kernel + a closed-form backward kernel.  Targets may be float32 (what ``train_mmwhs_noPad.py:181`` passes after
``.float()``) or the dataset's own bool / uint8 one-hot (``dataset/MMWHS.py:392,414-425``), which the kernels read as
bytes.
"""
import torch
import torch.nn as nn

from .. import ops


class MDiceLoss(nn.Module):
    _W = (0.7, 0.3)        # (dice weight, BCE weight)

    def __init__(self, do_sigmoid=True, process_group=None):
        super().__init__()
        if not do_sigmoid:
            raise NotImplementedError("micformer_b200: MDiceLoss(do_sigmoid=False) is not on the reference's path")
        self.do_sigmoid = do_sigmoid
        self.labels = ['backgroud', 'CT-A', 'CT-B', 'CT-C', 'CT-D', 'CT-E', 'CT-F', 'CT-G']
        self.device = "cpu"
        # None = per-process Dice sums (what the reference computes); a process group = global-batch Dice
        self.process_group = process_group

    @property
    def grad_reduce(self) -> str:
        """How data-parallel ranks must combine parameter gradients of this loss: "mean" for the reference's per-process
        loss, "sum" when ``process_group`` makes the loss global (its backward already is d(global loss)/d(local logits))."""
        return "sum" if self.process_group is not None else "mean"

    def forward(self, inputs, target):
        if target.dtype not in (torch.float32, torch.uint8, torch.bool):
            target = target.float()
        world = 1
        if self.process_group is not None:
            world = torch.distributed.get_world_size(self.process_group)
        return ops.DiceBceLossFn.apply(inputs.contiguous(), target.contiguous(), self.process_group, world, *self._W)

    def binary_dice(self, inputs, targets, label_index, metric_mode=False):
        """dice.py:130-151 -- validation metric helper (not on the training hot path; plain torch)."""
        smooth = 1.
        if self.do_sigmoid:
            inputs = torch.sigmoid(inputs)
        if metric_mode:
            inputs = inputs > 0.5
            if targets.sum() == 0:
                print(f"No {self.labels[label_index]} for this patient")
                return torch.tensor(1. if inputs.sum() == 0 else 0., device=inputs.device)
        intersection = torch.sum(inputs * targets)
        if metric_mode:
            return (2 * intersection) / ((inputs.sum() + targets.sum()) * 1.0)
        return 1 - (2 * intersection + smooth) / (inputs.pow(2).sum() + targets.pow(2).sum() + smooth)

    @staticmethod
    def compute_intersection(inputs, targets):
        return torch.sum(inputs * targets)

    def metric(self, inputs, target):
        """dice.py:168-175."""
        return [[self.binary_dice(inputs[j, i], target[j, i], i, True) for i in range(target.size(1))]
                for j in range(target.size(0))]


class MDiceLoss_Val(MDiceLoss):
    """dice.py:178-230: the validation criterion of ``train_mmwhs_noPad.py:109-110`` -- Dice term only."""
    _W = (1.0, 0.0)
