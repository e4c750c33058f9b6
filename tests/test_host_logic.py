"""Host-side logic that needs no GPU: window geometry (get_window_size clamp + trailing pad, reference M:135-145,
345-348), the gradient arena bookkeeping, and the reference arm of bench.py (contract keys of the JSON line)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_window_geometry_matches_reference_rules():
    from micformer_b200.ops import window_geometry
    # window 2 on an even grid: no clamp, no pad
    assert window_geometry((32, 32, 32), (2, 2, 2)) == ((2, 2, 2), (32, 32, 32))
    # window 7 at the four stages of a 128^3 volume (32->35, 16->21, 8->14, stage 3 clamps to 4^3)
    assert window_geometry((32, 32, 32), (7, 7, 7)) == ((7, 7, 7), (35, 35, 35))
    assert window_geometry((16, 16, 16), (7, 7, 7)) == ((7, 7, 7), (21, 21, 21))
    assert window_geometry((8, 8, 8), (7, 7, 7)) == ((7, 7, 7), (14, 14, 14))
    assert window_geometry((4, 4, 4), (7, 7, 7)) == ((4, 4, 4), (4, 4, 4))
    # anisotropic grid: each axis is clamped / padded on its own
    assert window_geometry((4, 9, 14), (7, 7, 7)) == ((4, 7, 7), (4, 14, 14))


def test_grad_arena_bookkeeping_cpu():
    from micformer_b200.arena import GradArena
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2, 2))]
    ps[1].requires_grad_(False)
    arena = GradArena(ps)
    assert arena.attached() and len(arena.params) == 2 and ps[1].grad is None
    assert arena.flat.numel() % 64 == 0 and all(p.grad.data_ptr() % 256 == arena.flat.data_ptr() % 256 for p in arena.params)
    (ps[0].sum() * 2 + ps[2].sum() * 3).backward()            # ordinary autograd accumulates in place into the slices
    assert arena.attached() and float(ps[0].grad.min()) == 2.0 and float(ps[2].grad.max()) == 3.0
    assert float(arena.flat.sum()) == 2.0 * 15 + 3.0 * 8
    ps[0].grad = None                                         # what optimizer.zero_grad(set_to_none=True) does
    assert not arena.attached()
    arena.zero()                                              # re-attaches and clears
    assert arena.attached() and float(arena.flat.abs().max()) == 0.0
    with pytest.raises(ValueError):
        GradArena([ps[1]])


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: the reference's CPU implementation (oracle port) timed on the host cores; rank 0 prints
    one JSON line with the metric / unit of the main arm, impl=reference, a cpu_baseline block and a zero-copy e2e block."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--size", "64", "--batch", "1"], capture_output=True, text=True, env=env, timeout=560)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "volumes/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("volumes/sec") and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank of a multi-process launch exits quietly
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--size", "64", "--batch", "1"], capture_output=True, text=True, env=dict(env, RANK="1"), timeout=120)
    assert out2.returncode == 0 and out2.stdout.strip() == ""


def test_fused_tail_algebra_design_artifact():
    """scripts/fused_tail_algebra.py: the composed ConvTranspose(k4,s4) o Conv3d(k3) weights planned for the decoder tail
     reproduce torch's two convolutions, including the bias at the volume border."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fused_tail_algebra as A
    for seed, dims in [(0, (3, 4, 5)), (1, (1, 2, 1)), (2, (2, 2, 2))]:
        err, nz = A.check(seed=seed, dims=dims)
        assert err < 1e-12 and nz == 216


def test_segmented_arena_for_model():
    """GradArena.for_model: [decoder + tail | encoder] segments partition the flat buffer; every parameter points into it"""
    from micformer_b200.arena import GradArena
    from micformer_b200.models.MICFormer_self import Head
    h = Head(embed_dim=24, num_classes=8)
    a = GradArena.for_model(h)
    (d0, d1), (e0, e1) = a.segments
    assert d0 == 0 and d1 == e0 and e1 == a.flat.numel() and a.attached()
    base = a.flat.data_ptr()
    for name, p in h.named_parameters():
        off = (p.grad.data_ptr() - base) // 4
        dec = name.startswith(("swin.up_layers.", "swin.concat_back_dim.", "swin.norm2.", "swin.reverse_patch_embedding.", "out_conv."))
        assert (d0 <= off < d1) if dec else (e0 <= off < e1), name
    assert sum(p.numel() for p in h.parameters()) <= a.flat.numel()
