"""GPU parity of the depth-wise causal conv1d kernels (csrc/causal_conv1d.cu, through the C ABI) against the reference's
own definition of the op, `self.act(self.conv1d(x)[..., :seqlen])` with a depth-wise nn.Conv1d(padding = k - 1)
(arm/Finetuning/mamba_simple.py:112-120, 673), evaluated by torch on the CPU in float64."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, b, silu):
    x64, w64 = x.double(), w.double()
    y = F.conv1d(x64, w64.unsqueeze(1), None if b is None else b.double(), padding=w.shape[1] - 1, groups=w.shape[0])[..., : x.shape[-1]]
    return F.silu(y) if silu else y


@pytest.mark.parametrize("shape", [(2, 8, 16), (3, 96, 196), (1, 5, 4), (2, 64, 1024), (2, 1536, 196), (2, 24, 197), (3, 7, 50)], ids=str)
@pytest.mark.parametrize("width", [4, 3, 2])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("silu,has_bias", [(True, True), (False, False)])
def test_causal_conv1d_parity(shape, width, dtype, silu, has_bias):
    from medical_image_analysis_b200.selective_scan_interface import causal_conv1d_fn
    B, D, L = shape
    g = torch.Generator().manual_seed(B * 1000 + D + L + width)
    x = torch.randn(B, D, L, generator=g).to(dtype)
    w = (torch.randn(D, width, generator=g) * 0.5)
    b = torch.randn(D, generator=g) if has_bias else None
    dy = torch.randn(B, D, L, generator=g).to(dtype)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    br = None if b is None else b.double().requires_grad_(True)
    yr = _ref(xr, wr, br, silu)
    yr.backward(dy.double())
    xg = x.cuda().requires_grad_(True)
    wg = w.cuda().requires_grad_(True)
    bg = None if b is None else b.cuda().requires_grad_(True)
    y = causal_conv1d_fn(xg, wg, bg, "silu" if silu else None)
    assert y.dtype == dtype and y.shape == x.shape
    y.backward(dy.cuda())
    rt, at = (1e-5, 1e-5) if dtype == torch.float32 else ((1e-2, 1e-2) if dtype == torch.bfloat16 else (2e-3, 2e-3))

    def close(got, ref, what, scale=1.0):
        got, ref = got.detach().double().cpu(), ref.detach()
        err = (got - ref).abs()
        tol = at * scale * max(1.0, ref.abs().max().item()) + rt * ref.abs()
        assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.3e} (ref max {ref.abs().max().item():.3e})"

    close(y, yr, "y")
    close(xg.grad, xr.grad, "dx")
    # weight grads are fp32 sums over batch * seqlen of products of (rounded) activations
    close(wg.grad, wr.grad, "dweight", 1.0 if dtype == torch.float32 else 0.1)
    if b is not None:
        close(bg.grad, br.grad, "dbias", 1.0 if dtype == torch.float32 else 0.1)


def test_causal_conv1d_on_xz_view_and_errors():
    """x is the first half of xz (mamba_simple.py:666): batch stride 2 d L, no copy needed; bad shapes raise."""
    from medical_image_analysis_b200.selective_scan_interface import causal_conv1d_fn
    torch.manual_seed(0)
    xz = torch.randn(2, 64, 100, device="cuda", dtype=torch.bfloat16)
    x = xz[:, :32]
    w = torch.randn(32, 4, device="cuda")
    y = causal_conv1d_fn(x, w, None, "silu")
    ref = _ref(x.cpu(), w.cpu(), None, True)
    assert torch.allclose(y.double().cpu(), ref, rtol=1e-2, atol=1e-2)
    with pytest.raises(RuntimeError):
        causal_conv1d_fn(torch.randn(2, 16, 12, device="cuda"), w, None, "silu")        # channel mismatch
    with pytest.raises(NotImplementedError):
        causal_conv1d_fn(x, w, None, "gelu")
