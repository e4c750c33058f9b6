"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what include/*.h
declares, the ctypes signatures agree with the header, and the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "micformer_b200.h")

_CT = {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "int64_t": ctypes.c_int64}


def _header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:const\s+char\s*\*|int64_t|int|void)\s+(mic_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(ctypes.c_void_p)
                else:
                    types.append(_CT[a.split()[-2] if len(a.split()) > 1 else a])
        decls[name] = types
    return decls


@pytest.fixture(scope="module")
def lib():
    from micformer_b200 import build, _native
    build.build()
    return _native.load()


def test_header_symbols_exported_and_signatures_match(lib):
    from micformer_b200 import _native
    decls = _header_decls()
    assert len(decls) >= 26
    for name, types in decls.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
        sig = _native.SIGNATURES[name]
        assert len(sig) == len(types), f"{name}: header has {len(types)} args, binding has {len(sig)}"
        for i, (a, b) in enumerate(zip(sig, types)):
            assert a is b, f"{name} arg {i}: binding {a} vs header {b}"
    assert set(_native.SIGNATURES) == set(decls), "binding and header disagree on the symbol set"


def test_version_and_error_string(lib):
    assert lib.mic_version() >= 100
    assert lib.mic_set_gemm_mode(7) != 0
    assert b"gemm mode" in lib.mic_last_error_string()
    assert lib.mic_set_gemm_mode(0) == 0


def test_argument_validation_without_gpu(lib):
    # pure host-side validation paths: no kernel is launched
    rc = lib.mic_window_attn_fwd(None, 0, None, None, 0, None, 0, None, 1, 2, 2, 2, 1, 16, 2, 2, 2, 1.0, None)
    assert rc != 0 and b"null" in lib.mic_last_error_string()
    rc = lib.mic_linear_fwd(None, 0, None, 0, 0, None, None, 0, 0, 0, 0, 0, None, 0, None, 0, None, 0, 0, None)
    assert rc != 0


def test_product_path_refuses_cpu_tensors():
    from micformer_b200.models import Head
    from micformer_b200.loss import MDiceLoss
    head = Head(embed_dim=24, num_classes=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        head(torch.randn(1, 2, 32, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        MDiceLoss()(torch.randn(1, 8, 4, 4, 4), torch.rand(1, 8, 4, 4, 4))
    with pytest.raises(ValueError):
        head(torch.randn(1, 3, 32, 32, 32))


def test_module_surface_matches_reference_structure():
    """Class names / constructor defaults / state_dict keys of the reference."""
    import inspect
    from micformer_b200.models import MICFormer_self as M
    from oracle import micformer_oracle as O
    sig = inspect.signature(M.Head.__init__)
    assert [p for p in sig.parameters][1:] == ["n_channels", "embed_dim", "num_classes", "window_size"]
    assert sig.parameters["embed_dim"].default == 96 and sig.parameters["num_classes"].default == 14
    msig = inspect.signature(M.MicFormer.__init__)
    assert msig.parameters["embed_dim"].default == 64 and msig.parameters["window_size"].default == (7, 7, 7)
    assert msig.parameters["drop_path_rate"].default == 0.2
    torch.manual_seed(0)
    head = M.Head(embed_dim=48, num_classes=8)
    sd = head.state_dict()
    shapes = O.param_shapes(O.TRAIN)
    assert list(sd.keys()) == list(shapes.keys())
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert abs(float(head.swin.patch_embed.proj.weight.detach().sum()) - (-1.4963787198)) < 1e-5   # same init stream
    # stochastic-depth ladder 0 .. 0.2 over 12 blocks, first block Identity (test.ipynb print(model))
    b0 = head.swin.layers[0].blocks1[0]
    assert isinstance(b0.drop_path, torch.nn.Identity)
    assert abs(head.swin.layers[3].blocks1[1].drop_path.drop_prob - 0.2) < 1e-6
    with pytest.raises(RuntimeError):      # MicFormer() own defaults are not runnable in the reference either (F9)
        M.MicFormer()(torch.zeros(1, 1, 32, 32, 32), torch.zeros(1, 1, 32, 32, 32))
