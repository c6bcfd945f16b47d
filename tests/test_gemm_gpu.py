"""tcgen05 GEMM (csrc/gemm_tcgen05.cu, through the C ABI) vs an fp64 torch contraction of the same bf16 / fp16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (M, N, K): tile-aligned, ragged M / N, K tail inside a 64-block, tiny, the shapes of the models
SHAPES = [(128, 128, 64), (256, 256, 256), (100, 72, 40), (1, 8, 8), (300, 56, 768), (392, 1536, 768), (12544, 768, 1536),
          (6400, 1024, 256), (400, 1024, 16384), (61, 3072, 1024), (197 * 4, 80, 768), (777, 333, 136)]


def _ref(a, w, bias, act):
    y = a.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        y = torch.nn.functional.silu(y)
    return y


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"M{s[0]}N{s[1]}K{s[2]}")
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_gemm_tn_parity(shape, dtype):
    from medical_image_analysis_b200.gemm import gemm_tn
    from tests.parity import cmp_f32, cmp_stored
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g) / K ** 0.5).to(dtype)
    w = torch.randn(N, K, generator=g).to(dtype)
    bias = torch.randn(N, generator=g)
    for act, has_bias, out_f32 in ((0, False, True), (0, True, False), (1, True, False), (2, True, True), (3, False, False)):
        ref = _ref(a, w, bias if has_bias else None, act)
        out = gemm_tn(a.cuda(), w.cuda(), bias.cuda() if has_bias else None, act, torch.float32 if out_f32 else None)
        tag = f"gemm M{M} N{N} K{K} act{act} bias{int(has_bias)} {'f32' if out_f32 else str(dtype)[6:]}"
        if out_f32:
            # fp32 accumulation of K products of bf16 / fp16 values in tensor memory vs fp64: a few 1e-6 relative to the
            # magnitude of the terms; tied to RMS(ref), rtol 1e-4 for the activation's approximations (erf / exp)
            cmp_f32(out, ref, tag, rtol=1e-4, atol_rms=1e-4)
        else:
            cmp_stored(out, ref, dtype, tag)


def test_gemm_strided_operands_and_errors():
    from medical_image_analysis_b200.gemm import gemm_tn
    g = torch.Generator().manual_seed(3)
    big = torch.randn(200, 512, generator=g).bfloat16().cuda()
    a = big[:, 128:384]                                    # row pitch 512, 16-byte aligned start
    w = torch.randn(96, 256, generator=g).bfloat16().cuda()
    out = gemm_tn(a, w, None, 0, torch.float32)
    ref = a.double() @ w.double().t()
    assert torch.allclose(out.double(), ref, rtol=1e-4, atol=1e-3)
    with pytest.raises(RuntimeError):
        gemm_tn(a.float(), w.float())
    with pytest.raises(RuntimeError):
        gemm_tn(a, w[:, :100])


@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_autograd_matches_torch(act):
    """LinearTC forward + the three backward GEMMs vs torch autograd of the same bf16 operands in fp32."""
    from medical_image_analysis_b200.gemm import linear
    from tests.parity import cmp_stored
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(3, 50, 136, generator=g) / 8).bfloat16()
    w = (torch.randn(200, 136, generator=g) / 8).bfloat16()
    b = torch.randn(200, generator=g)
    dy = torch.randn(3, 50, 200, generator=g).bfloat16()
    xg, wg, bg = x.cuda().requires_grad_(), w.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = linear(xg, wg, bg, act)
    y.backward(dy.cuda())
    xr, wr, br = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    yr = torch.nn.functional.linear(xr, wr, br)
    yr = {0: lambda t: t, 1: torch.relu, 2: torch.nn.functional.gelu}[act](yr)
    yr.backward(dy.double())
    # training-mode GELU keeps the bf16 pre-activation for the backward and applies GELU to it: two roundings
    cmp_stored(y, yr, torch.bfloat16, f"linear act{act} y", n_ulp=2.0 if act == 2 else 1.0)
    if act == 2:
        # the GELU derivative is evaluated at the bf16-ROUNDED pre-activation kept for the backward (2^-9 relative on its
        # argument) and dy is rounded to bf16 again after it: ~1 % of the gradient's scale, not ulp-level
        for got, ref, nm in ((xg.grad, xr.grad, "dx"), (wg.grad, wr.grad, "dw")):
            err = (got.double().cpu() - ref).abs()
            assert float(err.max()) < 3e-2 * float(ref.pow(2).mean().sqrt()) + 2e-2 * float(ref.abs().max()), (nm, float(err.max()))
    else:
        cmp_stored(xg.grad, xr.grad, torch.bfloat16, f"linear act{act} dx", n_ulp=2.0)      # dy rounded again after the activation derivative
        cmp_stored(wg.grad, wr.grad, torch.bfloat16, f"linear act{act} dw", n_ulp=2.0)
    assert torch.allclose(bg.grad.double().cpu(), br.grad, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("shape", [(128, 128, 64), (200, 136, 72), (392, 80, 768), (61, 3072, 1024), (1000, 56, 264), (64, 256, 4096)],
                         ids=lambda s: f"M{s[0]}N{s[1]}K{s[2]}")
@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, True), (True, False)])
def test_gemm_mn_major_operands(shape, a_mn, b_mn):
    """Operands read in place with their M / N index contiguous (UMMA a_major / b_major = MN): what the backward of a linear
    layer contracts, checked against fp64 on the same bf16 values."""
    from medical_image_analysis_b200.gemm import gemm
    from tests.parity import cmp_f32
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need a 16-byte row pitch")
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    a = (torch.randn(M, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, K, generator=g).bfloat16()
    ref = a.double() @ b.double().t()
    a_dev = a.t().contiguous().cuda() if a_mn else a.cuda()           # stored [K][M] when MN-major
    b_dev = b.t().contiguous().cuda() if b_mn else b.cuda()
    out = gemm(a_dev, b_dev, a_mn, b_mn, torch.float32)
    cmp_f32(out, ref, f"gemm mn a{int(a_mn)} b{int(b_mn)} M{M} N{N} K{K}", rtol=1e-4, atol_rms=1e-4)
