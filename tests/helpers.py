"""Shared test helpers: golden fixtures and seeded block cases (mirrors tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import micformer_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden():
    vec = np.load(os.path.join(GOLDEN_DIR, "golden_vectors.npz"))
    with open(os.path.join(GOLDEN_DIR, "golden_meta.json")) as f:
        meta = json.load(f)
    return vec, meta


def block_weights(C, cross, seed, prefix="blk"):
    """Same generator recipe as make_golden.block_weights; keys prefixed with ``prefix.``."""
    shapes = O._block_shapes("blk", C, cross, 16, 4.0)
    out = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed(O._key_seed(key, seed))
        if ".norm" in key and key.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in key:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        out[prefix + key[len("blk"):]] = t
    return out


def block_inputs(C, dims, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, *dims, C, generator=g)
    xa = torch.randn(1, *dims, C, generator=g)
    gy = torch.randn(1, *dims, C, generator=g)
    return x, xa, gy


def rel_err(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
