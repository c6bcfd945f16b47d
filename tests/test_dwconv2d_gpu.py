"""GPU parity of the depth-wise 3x3 conv2d (+ bias, + SiLU) kernels (csrc/dwconv2d.cu, through the C ABI) against the
reference's own definition, `self.act(self.conv2d(x))` with nn.Conv2d(C, C, 3, padding=1, groups=C)
(R2GenCSR/VMamba/classification/models/vmamba.py:574-582, 1120-1122), evaluated by torch on the CPU in float64."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 8, 14, 14), (3, 96, 7, 7), (1, 5, 1, 1), (2, 16, 28, 28), (1, 4, 80, 80), (2, 768, 14, 14), (2, 6, 3, 17)], ids=str)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=["f32", "bf16", "f16"])
@pytest.mark.parametrize("silu,has_bias", [(True, True), (False, False), (True, False)])
def test_dwconv2d_parity(shape, dtype, silu, has_bias):
    from medical_image_analysis_b200.vmamba import DWConv2dFn
    B, C, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + C + H * W)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    w = torch.randn(C, 1, 3, 3, generator=g) * 0.4
    b = torch.randn(C, generator=g) if has_bias else None
    dy = torch.randn(B, C, H, W, generator=g).to(dtype)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    br = None if b is None else b.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, padding=1, groups=C)
    if silu:
        yr = F.silu(yr)
    yr.backward(dy.double())
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    bg = None if b is None else b.cuda().requires_grad_(True)
    y = DWConv2dFn.apply(xg, wg, bg, silu)
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
    close(wg.grad, wr.grad, "dweight", 1.0 if dtype == torch.float32 else 0.1)
    if b is not None:
        close(bg.grad, br.grad, "dbias", 1.0 if dtype == torch.float32 else 0.1)


def test_ss2d_uses_the_fused_conv_and_matches_torch():
    """SS2D.forward with the fused conv + SiLU equals the same module evaluated with nn.Conv2d + nn.SiLU (fwd + grads)."""
    from medical_image_analysis_b200 import vmamba
    torch.manual_seed(0)
    m = vmamba.SS2D(d_model=32, d_state=1, ssm_ratio=2.0, dt_rank="auto", d_conv=3, forward_type="v2").cuda()
    x = torch.randn(2, 14, 14, 32, device="cuda", requires_grad=True)
    y = m(x)
    y.square().mean().backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    gx1 = x.grad.clone()
    m.zero_grad(); x.grad = None
    orig = vmamba.dwconv2d_silu
    vmamba.dwconv2d_silu = lambda t, conv, silu=True: F.silu(conv(t))
    try:
        y2 = m(x)
        y2.square().mean().backward()
    finally:
        vmamba.dwconv2d_silu = orig
    assert torch.allclose(y, y2, rtol=1e-4, atol=1e-5)
    assert torch.allclose(gx1, x.grad, rtol=1e-3, atol=1e-5)
    for a, p in zip(g1, m.parameters()):
        assert torch.allclose(a, p.grad, rtol=2e-3, atol=1e-5), (a - p.grad).abs().max()
