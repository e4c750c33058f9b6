"""Validation-time callers (micformer_b200/inference.py) on CPU with stand-in predictors: the window
lattice and blending of the sliding-window restatement, and the reference's mean-Dice metric."""
import pytest
import torch

from micformer_b200 import inference as I


def test_window_lattice_matches_the_reference_call():
    # roi 128, overlap 0.5 -> step 64; a 128^3 volume is one window, 160 needs 2 per axis with the last clamped to 32
    assert I.window_lattice((128, 128, 128), (128, 128, 128), 0.5) == [(0, 0, 0)]
    assert I._scan_starts(160, 128, 0.5) == [0, 32]
    assert I._scan_starts(192, 128, 0.5) == [0, 64]
    assert I._scan_starts(200, 128, 0.5) == [0, 64, 72]
    lat = I.window_lattice((192, 128, 160), (128, 128, 128), 0.5)
    assert lat == [(0, 0, 0), (0, 0, 32), (64, 0, 0), (64, 0, 32)]              # last axis fastest
    assert I._scan_starts(9, 8, 0.95) == [0, 1]                                  # step never drops below 1


def test_single_window_is_the_predictor_itself():
    calls = []

    def pred(w):
        calls.append(tuple(w.shape))
        return torch.cat([w * 2, w + 1], 1)
    x = torch.randn(2, 1, 8, 8, 8)
    y = I.sliding_window_inference(x, (8, 8, 8), 1, pred, overlap=0.5)
    assert torch.equal(y, torch.cat([x * 2, x + 1], 1)) and calls == [(1, 1, 8, 8, 8)] * 2


@pytest.mark.parametrize("size,roi,overlap,sw", [((12, 8, 10), (8, 8, 8), 0.5, 1), ((9, 13, 8), (4, 6, 8), 0.25, 3),
                                                ((5, 6, 7), (8, 8, 8), 0.5, 2)])
def test_blending_is_a_partition_of_unity(size, roi, overlap, sw):
    """an identity predictor must give the input back whatever the overlap pattern (also through the small-volume pad)"""
    x = torch.randn(2, 3, *size)
    y = I.sliding_window_inference(x, roi, sw, lambda w: w.clone(), overlap=overlap)
    assert y.shape == x.shape and float((y - x).abs().max()) < 1e-6


def test_overlapping_predictions_are_averaged():
    # predictor returns a constant that identifies the call; voxels covered by windows k and m must hold their mean
    counter = {"n": 0}

    def pred(w):
        counter["n"] += 1
        return torch.full_like(w, float(counter["n"]))
    x = torch.zeros(1, 1, 4, 4, 12)
    y = I.sliding_window_inference(x, (4, 4, 8), 1, pred, overlap=0.5)          # starts along the last axis: 0, 4
    assert counter["n"] == 2
    assert float(y[0, 0, 0, 0, 0]) == 1.0 and float(y[0, 0, 0, 0, 6]) == 1.5 and float(y[0, 0, 0, 0, 11]) == 2.0


def _meandice_loop(pred, label, num_class):
    """the reference's definition, written out class by class (train_mmwhs_noPad.py:392-407)"""
    total = 0.0
    for c in range(1, num_class):
        pb, lb = (pred == c).double(), (label == c).double()
        total += (2.0 * (pb * lb).sum() + 1e-6) / (pb.sum() + lb.sum() + 1e-6)
    return total / (num_class - 1)


def test_meandice_equals_the_per_class_definition():
    g = torch.Generator().manual_seed(0)
    for shape, nc in [((2, 6, 7, 5), 8), ((1, 9, 9, 9), 3)]:
        pred = torch.randint(0, nc, shape, generator=g)
        label = torch.randint(0, nc, shape, generator=g)
        assert abs(float(I.meandice(pred, label, nc)) - float(_meandice_loop(pred, label, nc))) < 1e-12
    # a class absent from both maps scores 1 through the smooth term; identical maps score 1
    pred = torch.zeros(1, 4, 4, 4, dtype=torch.long); pred[0, 0] = 1
    assert abs(float(I.meandice(pred, pred, 4)) - 1.0) < 1e-9
    lab = torch.zeros_like(pred); lab[0, 1] = 1
    assert abs(float(I.meandice(pred, lab, 4)) - float(_meandice_loop(pred, lab, 4))) < 1e-12


def test_evaluate_line_of_the_validation_loop():
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(1, 8, 6, 6, 6, generator=g)
    cls = torch.randint(0, 8, (1, 6, 6, 6), generator=g)
    onehot = torch.nn.functional.one_hot(cls, 8).permute(0, 4, 1, 2, 3).bool()
    ref = _meandice_loop(torch.argmax(torch.softmax(logits, 1), 1), cls, 8)
    assert abs(float(I.evaluate(logits, onehot, 8)) - float(ref)) < 1e-12
