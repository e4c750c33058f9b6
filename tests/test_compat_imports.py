"""The drop-in import surface: with ``compat/`` first on ``sys.path`` the reference script's own import lines
(MicFormer/train_mmwhs_noPad.py:19-20,26 and test.ipynb cell 0) resolve to micformer_b200, the state_dict interchanges with
the unmodified reference in both directions, and FusedAdam interchanges optimizer checkpoints with torch.optim.Adam."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

IMPORT_LINES = """
import sys
sys.path.insert(0, {compat!r}); sys.path.insert(1, {root!r})
from loss import MDiceLoss                                   # train_mmwhs_noPad.py:19
from loss.dice import MDiceLoss_Val                          # train_mmwhs_noPad.py:20
from MMWHS_pre.Multi_modal.SymCFNet.models.MICFormer_self import Head      # train_mmwhs_noPad.py:26
from models.MICFormer_self import Head as Head2, CrossTransformerBlock3D, MicFormer   # test.ipynb cell 0, M:1058-1063
from models.STN import SpatialTransformer, Re_SpatialTransformer
import micformer_b200.models.MICFormer_self as M
assert Head is M.Head and Head2 is M.Head and MicFormer is M.MicFormer
model = Head(embed_dim=48, num_classes=8)                    # train_mmwhs_noPad.py:92
crit, crit_val = MDiceLoss(), MDiceLoss_Val()                # :108-110
metric = crit_val.metric
n = sum(p.numel() for p in model.parameters() if p.requires_grad)        # utils.count_parameters
assert n == 61722608, n                                      # SURVEY F7
assert len(model.state_dict()) == 1626
print("ok")
"""


def test_reference_script_import_lines_resolve_through_compat():
    code = IMPORT_LINES.format(compat=os.path.join(ROOT, "compat"), root=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.reference
def test_state_dict_interchanges_with_the_unmodified_reference():
    from _reference_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("/root/reference not present")
    ref_models, _ = load_reference()
    from micformer_b200.models.MICFormer_self import Head
    torch.manual_seed(0)
    ref = ref_models.Head(embed_dim=24, num_classes=8)
    torch.manual_seed(0)
    ours = Head(embed_dim=24, num_classes=8)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())                    # same keys, same order
    for k in sd_ref:                                                      # same init stream under the same seed
        assert sd_ref[k].shape == sd_ours[k].shape and torch.equal(sd_ref[k], sd_ours[k]), k
    # reference checkpoint -> ours (utils.py:125-138 reload_ckpt_bis) and ours -> reference, strict
    for v in sd_ref.values():
        v.add_(0.125)
    ours.load_state_dict(sd_ref, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), sd_ref.values()))
    ref.load_state_dict(ours.state_dict(), strict=True)


def test_mdiceloss_val_surface():
    from micformer_b200.loss import MDiceLoss, MDiceLoss_Val
    v = MDiceLoss_Val()
    assert isinstance(v, torch.nn.Module) and v._W == (1.0, 0.0) and MDiceLoss()._W == (0.7, 0.3)
    assert v.labels[0] == "backgroud" and callable(v.metric) and callable(v.binary_dice)
    # metric path is plain torch (validation only): runs on CPU exactly like loss/dice.py:168-175
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 4, 4, 4, generator=g)
    t = (torch.rand(2, 3, 4, 4, 4, generator=g) > 0.5).float()
    d = v.metric(x, t)
    p = (torch.sigmoid(x[1, 2]) > 0.5).float()
    assert abs(float(d[1][2]) - float(2 * (p * t[1, 2]).sum() / (p.sum() + t[1, 2].sum()))) < 1e-6
    with pytest.raises(RuntimeError):
        v(x, t)           # the loss itself runs on the sm_100a kernels only: a CPU tensor raises, no fallback
