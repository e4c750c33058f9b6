"""micformer_b200/data.py on CPU: the sample contract of dataset/MMWHS.py and the transform chain of
train_mmwhs_noPad.py:116-130 against an independent numpy restatement of the documented MONAI semantics."""
import numpy as np
import torch

from micformer_b200 import data as D


def _normalize_np(img):
    out = img.astype(np.float32).copy()
    for c in range(out.shape[0]):
        m = out[c] != 0
        if not m.any():
            continue
        mu, sd = out[c][m].mean(), out[c][m].std()
        out[c][m] = (out[c][m] - mu) / (sd if sd != 0 else 1.0)
    return out


def test_sample_contract():
    ds = D.SyntheticMMWHS(n=3, size=16, seed=1)
    s = ds[2]
    assert len(ds) == 3 and set(s) >= {"patient_id", "image", "label", "seg_path", "crop_indexes", "et_present", "supervised"}
    assert s["image"].shape == (2, 16, 16, 16) and s["image"].dtype == torch.float16
    assert s["label"].shape == (8, 16, 16, 16) and s["label"].dtype == torch.bool
    assert bool((s["label"].sum(0) == 1).all())                       # one-hot over the 8 planes
    assert float(s["image"][:, 0].abs().max()) == 0.0                 # zero background shell
    assert torch.equal(ds[2]["image"], s["image"]) and not torch.equal(ds[1]["image"], s["image"])


def test_normalize_intensity_matches_numpy_restatement():
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 6, 7, 5, generator=g) * 3 + 1
    img[0, :2] = 0                                                    # zeros stay zero and do not enter the statistics
    img[1] = 0                                                        # an all-zero channel is returned unchanged
    out = D.normalize_intensity_nonzero_channelwise(img)
    ref = _normalize_np(img.numpy())
    assert float(np.abs(out.numpy() - ref).max()) < 1e-5
    assert float(out[0, :2].abs().max()) == 0.0 and float(out[1].abs().max()) == 0.0
    nz = out[0][img[0] != 0]
    assert abs(float(nz.mean())) < 1e-5 and abs(float(nz.std(unbiased=False)) - 1.0) < 1e-4
    # batched input: every (sample, channel) is normalised on its own
    both = D.normalize_intensity_nonzero_channelwise(torch.stack([img, img * 2 + 5 * (img != 0)]))
    assert float((both[0] - out).abs().max()) < 1e-6 and float((both[1, 0] - out[0]).abs().max()) < 1e-4
    # constant non-zero channel: std 0 -> divide by 1
    const = torch.full((1, 2, 2, 2), 3.0)
    assert float(D.normalize_intensity_nonzero_channelwise(const).abs().max()) == 0.0


def test_train_transform_chain():
    ds = D.SyntheticMMWHS(n=1, size=8, seed=3)
    s = ds[0]
    g = torch.Generator().manual_seed(7)
    out = D.train_transform(s, g)
    # replay the same draws: three flip decisions, one scale factor, one shift offset
    g2 = torch.Generator().manual_seed(7)
    img, lab = s["image"].numpy().astype(np.float32), s["label"].numpy()
    for axis in (0, 1, 2):
        if float(torch.rand((), generator=g2)) < 0.5:
            img, lab = np.flip(img, axis + 1), np.flip(lab, axis + 1)
    ref = _normalize_np(np.ascontiguousarray(img))
    ref = ref * (1.0 + (float(torch.rand((), generator=g2)) * 0.2 - 0.1))
    ref = ref + (float(torch.rand((), generator=g2)) * 0.2 - 0.1)
    assert out["image"].dtype == torch.float32 and float(np.abs(out["image"].numpy() - ref).max()) < 1e-5
    assert np.array_equal(out["label"].numpy(), lab) and out["patient_id"] == s["patient_id"]
    # image and label are flipped together: wherever the (normalised, scaled, shifted) image is background, the label is class 0
    shift = float(out["image"][0, 0, 0, 0])                            # a shell voxel: 0 * (1 + f) + offset
    assert bool(out["label"][0][out["image"][0] == shift].all())
    val = D.val_transform(s)
    assert float(np.abs(val["image"].numpy() - _normalize_np(s["image"].numpy().astype(np.float32))).max()) < 1e-5
    assert torch.equal(val["label"], s["label"])


def test_dataloader_batches_feed_the_training_loop_contract():
    ds = D.SyntheticMMWHS(n=4, size=8, seed=5, transform=D.val_transform)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)
    batch = next(iter(loader))
    x, lab = batch["image"].float(), batch["label"].float()          # train_mmwhs_noPad.py:177
    assert x.shape == (2, 2, 8, 8, 8) and lab.shape == (2, 8, 8, 8, 8) and len(batch["patient_id"]) == 2
