"""GPU parity of the reference-facing Python surface: autograd Functions, module mirrors, mamba_ssm shim."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(seed, batch, dim, L, N, G, dtype, has_z=False):
    g = torch.Generator().manual_seed(seed)
    A = -0.5 * torch.rand(dim, N, generator=g)
    B = torch.randn(batch, G, N, L, generator=g).to(dtype)
    C = torch.randn(batch, G, N, L, generator=g).to(dtype)
    D = torch.randn(dim, generator=g)
    bias = 0.5 * torch.rand(dim, generator=g)
    u = torch.randn(batch, dim, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(batch, dim, L, generator=g)).to(dtype)
    z = torch.randn(batch, dim, L, generator=g).to(dtype) if has_z else None
    dout = torch.randn(batch, dim, L, generator=g)
    return dict(u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, bias=bias, dout=dout)


def _leafs(d, dev):
    return {k: (None if v is None else v.to(dev).detach().clone().requires_grad_(k != "dout")) for k, v in d.items()}


def _close(a, b, rtol, atol, what):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    assert bool((err <= atol * max(1.0, b.abs().max().item()) + rtol * b.abs()).all()), f"{what}: max err {err.max().item():.3e}"


@pytest.mark.parametrize("fn_name", ["SelectiveScanOflex", "SelectiveScanCore", "SelectiveScanMamba"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_vmamba_autograd_functions(fn_name, dtype):
    """vmamba.py:250-312 call pattern: Fn.apply(u, delta, A, B, C, D, delta_bias, softplus, nrows, backnrows, oflex)."""
    from medical_image_analysis_b200 import selective_scan_interface as ssi
    from oracle.selective_scan_ref import selective_scan_ref
    fn = getattr(ssi, fn_name)
    d = _mk(0, 2, 24, 130, 2, 2, dtype)
    g = _leafs(d, "cuda")
    out = fn.apply(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["bias"], True, 1, 1, True)
    assert out.dtype == (torch.float32 if fn_name == "SelectiveScanOflex" else dtype)      # oflex returns fp32
    out.backward(g["dout"].to(out.dtype))
    c = _leafs(d, "cpu")
    ref = selective_scan_ref(c["u"].float(), c["delta"].float(), c["A"], c["B"].float(), c["C"].float(), c["D"], None, c["bias"], True)
    ref.backward(c["dout"])
    lowp = dtype != torch.float32
    _close(out, ref, *( (1e-2, 1e-2) if (lowp and fn_name != "SelectiveScanOflex") else (1e-4, 1e-4)), "out")
    for k in ("u", "delta", "A", "B", "C", "D", "bias"):
        _close(g[k].grad, c[k].grad, *((2e-2, 2e-2) if lowp else (2e-4, 2e-4)), f"d{k}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_selective_scan_fn_with_z_and_last_state(dtype):
    """mamba_ssm API as the ARM mixer calls it (mamba_simple.py:693-704): 3-D B/C, z gate, last state."""
    from medical_image_analysis_b200.selective_scan_interface import selective_scan_fn
    from oracle.selective_scan_ref import selective_scan_ref
    d = _mk(1, 2, 16, 197, 16, 1, dtype, has_z=True)
    d["B"], d["C"] = d["B"].squeeze(1), d["C"].squeeze(1)
    g, c = _leafs(d, "cuda"), _leafs(d, "cpu")
    out, last = selective_scan_fn(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], z=g["z"], delta_bias=g["bias"],
                                  delta_softplus=True, return_last_state=True)
    assert out.dtype == dtype and tuple(last.shape) == (2, 16, 16)
    out.backward(g["dout"].to(dtype))
    f = lambda t: t.float() if t is not None else None
    ref, rlast = selective_scan_ref(f(c["u"]), f(c["delta"]), c["A"], f(c["B"]), f(c["C"]), c["D"], f(c["z"]), c["bias"], True, True)
    ref.backward(c["dout"].to(dtype).float())
    lowp = dtype != torch.float32
    _close(out, ref, *((1e-2, 1e-2) if lowp else (1e-4, 1e-4)), "out")
    _close(last, rlast, 1e-4, 1e-4, "last_state")
    for k in ("u", "delta", "A", "B", "C", "D", "z", "bias"):
        assert g[k].grad.shape == c[k].grad.shape, k
        _close(g[k].grad, c[k].grad, *((3e-2, 3e-2) if lowp else (3e-4, 3e-4)), f"d{k}")


def test_module_mirrors_signatures():
    """selective_scan_cuda_oflex.fwd/bwd positional calls exactly as vmamba.py:299, 309 and the reference test make them."""
    import medical_image_analysis_b200.dropin as dropin
    dropin.install(force=True)
    import selective_scan_cuda
    import selective_scan_cuda_core
    import selective_scan_cuda_oflex
    d = _mk(2, 2, 8, 300, 1, 2, torch.bfloat16)
    g = {k: (None if v is None else v.cuda()) for k, v in d.items()}
    out, x = selective_scan_cuda_oflex.fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["bias"], True, 1, True)
    assert out.dtype == torch.float32 and tuple(x.shape) == (2, 8, 2, 2) and x.dtype == torch.float32
    res = selective_scan_cuda_oflex.bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["bias"], g["dout"].cuda(), x, True, 1)
    assert len(res) == 7 and res[0].dtype == torch.bfloat16 and res[2].dtype == torch.float32 and res[3].dtype == torch.bfloat16
    out2, x2 = selective_scan_cuda_core.fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], None, None, False, 1)
    assert out2.dtype == torch.bfloat16
    r2 = selective_scan_cuda_core.bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], None, None, g["dout"].bfloat16(), x2, False, 1)
    assert r2[5] is None and r2[6] is None
    r3 = selective_scan_cuda.fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], None, g["bias"], True)
    assert len(r3) == 2
    zz = torch.randn_like(g["u"])
    r4 = selective_scan_cuda.fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], zz, g["bias"], True)
    assert len(r4) == 3
    r5 = selective_scan_cuda.bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], zz, g["bias"], g["dout"].bfloat16(), r4[1], r4[0],
                                 None, True, True)
    assert len(r5) == 9


def test_reference_grid_case_delta_groups():
    """One cell of the reference's own parity grid (test_selective_scan.py:364-393): dim 768, delta rows 24, d_state 1,
    two B/C groups, bf16, seqlen 512, with the oflex delta-group semantics (:453-457, :510-517)."""
    from medical_image_analysis_b200 import selective_scan_cuda_oflex as oflex
    from oracle import ss_ref_c
    torch.manual_seed(0)
    batch, dim, dim1, L, N, G = 2, 768, 24, 512, 1, 2
    A = -0.5 * torch.rand(dim, N)
    B = torch.randn(batch, G, N, L).bfloat16()
    C = torch.randn(batch, G, N, L).bfloat16()
    D = torch.randn(dim)
    bias = 0.5 * torch.rand(dim1)
    u = torch.randn(batch, dim, L).bfloat16()
    delta = (0.5 * torch.rand(batch, dim1, L)).bfloat16()
    dout = torch.randn(batch, dim, L)
    dev = [t.cuda() for t in (u, delta, A, B, C, D, bias)]
    out, x = oflex.fwd(*dev, True, 1, True)
    r_out, _, r_last = ss_ref_c.fwd(u, delta, A, B, C, D, None, bias, True)
    _close(out, r_out, 1e-4, 1e-4, "out")
    _close(x[:, :, -1, 1::2], r_last, 1e-4, 1e-4, "last_state")
    du, dd, dA, dB, dC, dD, db = oflex.bwd(*dev, dout.cuda(), x, True, 1)
    ref = ss_ref_c.bwd(u, delta, A, B, C, D, None, bias, dout, True)
    assert tuple(dd.shape) == (batch, dim1, L) and tuple(db.shape) == (dim1,)
    for name, got in (("du", du), ("ddelta", dd), ("dB", dB), ("dC", dC)):
        _close(got, ref[name], 1e-2, 1e-2, name)
    for name, got in (("dA", dA), ("dD", dD), ("ddelta_bias", db)):
        _close(got, ref[name], 1e-4, 1e-4, name)


@pytest.mark.parametrize("with_out_proj", [False, True])
def test_mamba_inner_fn_matches_reference_slow_path(with_out_proj):
    """mamba_inner_fn(_no_out_proj) vs the reference's own slow path (mamba_simple.py:665-709), fwd + all grads."""
    from medical_image_analysis_b200.selective_scan_interface import mamba_inner_fn, mamba_inner_fn_no_out_proj
    from oracle.mamba_inner_ref import mamba_inner_ref
    torch.manual_seed(3)
    b, d_model, d_inner, N, R, L, W = 2, 12, 24, 16, 4, 50, 4
    P = dict(xz=torch.randn(b, 2 * d_inner, L), conv_w=torch.randn(d_inner, 1, W) * 0.3, conv_b=torch.randn(d_inner) * 0.1,
             x_proj=torch.randn(R + 2 * N, d_inner) * 0.2, dt_proj=torch.randn(d_inner, R) * 0.5,
             A=-torch.rand(d_inner, N) - 0.1, D=torch.randn(d_inner), bias=torch.rand(d_inner) - 3.0,
             out_w=torch.randn(d_model, d_inner) * 0.2, out_b=torch.randn(d_model) * 0.1)
    g = {k: v.cuda().requires_grad_() for k, v in P.items()}
    c = {k: v.clone().requires_grad_() for k, v in P.items()}
    if with_out_proj:
        out = mamba_inner_fn(g["xz"], g["conv_w"], g["conv_b"], g["x_proj"], g["dt_proj"], g["out_w"], g["out_b"], g["A"], None, None,
                             g["D"], delta_bias=g["bias"], delta_softplus=True)
        ref = mamba_inner_ref(c["xz"], c["conv_w"], c["conv_b"], c["x_proj"], c["dt_proj"], c["A"], c["D"], c["bias"], c["out_w"], c["out_b"])
    else:
        out = mamba_inner_fn_no_out_proj(g["xz"], g["conv_w"], g["conv_b"], g["x_proj"], g["dt_proj"], g["A"], None, None, g["D"],
                                         delta_bias=g["bias"], delta_softplus=True)
        ref = mamba_inner_ref(c["xz"], c["conv_w"], c["conv_b"], c["x_proj"], c["dt_proj"], c["A"], c["D"], c["bias"])
    w = torch.randn(ref.shape)
    (out * w.cuda()).sum().backward()
    (ref * w).sum().backward()
    _close(out, ref, 2e-4, 2e-4, "out")
    keys = list(P) if with_out_proj else [k for k in P if not k.startswith("out_")]
    for k in keys:
        _close(g[k].grad, c[k].grad, 1e-3, 1e-3, f"d{k}")


def test_bimamba_inner_fn_is_two_directions():
    from medical_image_analysis_b200.selective_scan_interface import bimamba_inner_fn, mamba_inner_fn_no_out_proj
    import torch.nn.functional as F
    torch.manual_seed(4)
    b, d_inner, N, R, L = 1, 8, 4, 2, 33
    xz = torch.randn(b, 2 * d_inner, L, device="cuda")
    cw, cb = torch.randn(d_inner, 1, 4, device="cuda") * 0.3, torch.zeros(d_inner, device="cuda")
    xp, dp = torch.randn(R + 2 * N, d_inner, device="cuda") * 0.2, torch.randn(d_inner, R, device="cuda")
    A, Ab = -torch.rand(d_inner, N, device="cuda"), -torch.rand(d_inner, N, device="cuda")
    D, bias = torch.randn(d_inner, device="cuda"), torch.rand(d_inner, device="cuda")
    ow = torch.randn(5, d_inner, device="cuda")
    out = bimamba_inner_fn(xz, cw, cb, xp, dp, ow, None, A, Ab, None, None, D, delta_bias=bias, delta_softplus=True)
    assert tuple(out.shape) == (b, L, 5) and torch.isfinite(out).all()
    fwd_only = F.linear(mamba_inner_fn_no_out_proj(xz, cw, cb, xp, dp, A, None, None, D, delta_bias=bias).transpose(1, 2), ow)
    assert not torch.allclose(out, fwd_only)


def test_selective_scan_cuda_bwd_recompute_out_z_is_a_kernel():
    """selective_scan_cuda.bwd(..., recompute_out_z=True) returns out * silu(z) (mia_silu_gate), equal to the forward's out_z."""
    from medical_image_analysis_b200 import selective_scan_cuda as ssc
    g = torch.Generator().manual_seed(4)
    u = torch.randn(2, 32, 40, generator=g).bfloat16().cuda()
    delta = (0.5 * torch.rand(2, 32, 40, generator=g)).bfloat16().cuda()
    A = (-0.5 * torch.rand(32, 4, generator=g)).cuda()
    B = torch.randn(2, 4, 40, generator=g).bfloat16().cuda()
    C = torch.randn(2, 4, 40, generator=g).bfloat16().cuda()
    z = torch.randn(2, 32, 40, generator=g).bfloat16().cuda()
    out, x, out_z = ssc.fwd(u, delta, A, B, C, None, z, None, True)
    res = ssc.bwd(u, delta, A, B, C, None, z, None, torch.randn_like(out), x, out, None, True, True)
    assert len(res) == 9
    ref = (out.float() * torch.nn.functional.silu(z.float())).to(out.dtype)
    assert torch.allclose(res[-1].float(), ref.float(), rtol=1e-2, atol=1e-2)
