"""Fused tcgen05 block kernels (csrc/block_mlp*.cu, block_attn*.cu) through the C ABI vs fp64 torch math of the
reference ops (Mlp M:28-34 inside the residual M:403-404,419-424) and vs the exact-fp32 unfused kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tensor_core_mode():
    """the fused kernels belong to gemm mode 1; the mode is process-global, so restore it for the other test modules"""
    from micformer_b200 import _native as N
    prev = N.get_gemm_mode()
    N.set_gemm_mode(1)
    yield
    N.set_gemm_mode(prev)


def _mlp_case(C, T, seed, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, C, generator=g)
    p = dict(gamma=1.0 + 0.1 * torch.randn(C, generator=g), beta=0.1 * torch.randn(C, generator=g),
             w1=(torch.rand(4 * C, C, generator=g) * 2 - 1) * 0.3, b1=0.1 * torch.randn(4 * C, generator=g),
             w2=(torch.rand(C, 4 * C, generator=g) * 2 - 1) * 0.3, b2=0.1 * torch.randn(C, generator=g))
    return x.to(dev), {k: v.to(dev) for k, v in p.items()}


def _mlp_ref64(x, p, rowscale=None, rps=1):
    x = x.double()
    xn = torch.nn.functional.layer_norm(x, (x.shape[-1],), p["gamma"].double(), p["beta"].double(), 1e-5)
    h = torch.nn.functional.gelu(xn @ p["w1"].double().t() + p["b1"].double())
    o = h @ p["w2"].double().t() + p["b2"].double()
    if rowscale is not None:
        o = o * rowscale.double().repeat_interleave(rps)[:, None]
    return x + o


@pytest.mark.parametrize("C,T", [(48, 128), (48, 1000), (24, 4096), (48, 65536)])
def test_fused_mlp_forward(C, T):
    from micformer_b200 import fused, _native as N
    dev = torch.device("cuda")
    N.set_gemm_mode(1)
    x, p = _mlp_case(C, T, 11 + C + T, dev)
    img = fused.mlp_images(p["w1"], p["w2"])
    img.refresh()
    y = fused.mlp_block_fwd(x, img, p["gamma"], p["beta"], p["b1"], p["b2"], None, 1, 1e-5)
    ref = _mlp_ref64(x, p)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err                      # split-bf16 products: ~2^-17 per term
    # per-sample DropPath scale on the branch (rows_per_sample = T/2)
    if T % 2 == 0:
        rs = torch.tensor([0.0, 1.25], device=dev)
        y2 = fused.mlp_block_fwd(x, img, p["gamma"], p["beta"], p["b1"], p["b2"], rs, T // 2, 1e-5)
        ref2 = _mlp_ref64(x, p, rs, T // 2)
        assert float((y2.double() - ref2).abs().max() / ref2.abs().max()) < 2e-5
        assert torch.equal(y2[: T // 2], x[: T // 2])          # dropped sample: the branch contributes exactly 0


@pytest.mark.parametrize("C,T", [(48, 128), (48, 1000), (24, 4096), (48, 65536)])
def test_fused_mlp_backward(C, T):
    from micformer_b200 import fused, _native as N
    dev = torch.device("cuda")
    N.set_gemm_mode(1)
    x, p = _mlp_case(C, T, 5 + C + T, dev)
    g = torch.Generator().manual_seed(99)
    dy = torch.randn(T, C, generator=g).to(dev)
    rs = torch.tensor([0.5, 1.25], device=dev) if T % 2 == 0 else None
    rps = T // 2 if rs is not None else 1
    img = fused.mlp_images(p["w1"], p["w2"])
    img.refresh()
    grads = {k: torch.zeros_like(p[k]) for k in ("gamma", "beta", "w1", "b1", "w2", "b2")}
    dx = fused.mlp_block_bwd(dy, x, img, p["gamma"], p["beta"], p["b1"], rs, rps, 1e-5, grads["gamma"], grads["beta"],
                             grads["w1"], grads["b1"], grads["w2"], grads["b2"])
    # fp64 autograd of the reference ops
    x64 = x.double().requires_grad_(True)
    p64 = {k: v.double().requires_grad_(True) for k, v in p.items()}
    y = _mlp_ref64(x64, p64, rs, rps)
    y.backward(dy.double())
    def rel(a, b):
        return float((a.double() - b).norm() / (b.norm() + 1e-30))
    assert rel(dx, x64.grad) < 2e-5, rel(dx, x64.grad)
    for k in grads:
        assert rel(grads[k], p64[k].grad) < 3e-5, (k, rel(grads[k], p64[k].grad))
    # accumulate semantics: a second call doubles the parameter gradients
    fused.mlp_block_bwd(dy, x, img, p["gamma"], p["beta"], p["b1"], rs, rps, 1e-5, grads["gamma"], grads["beta"],
                        grads["w1"], grads["b1"], grads["w2"], grads["b2"])
    assert rel(grads["w1"], 2 * p64["w1"].grad) < 3e-5
