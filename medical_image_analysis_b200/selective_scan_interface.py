"""Autograd surface of the selective scan, mirroring the reference's.

* ``SelectiveScanOflex`` / ``SelectiveScanCore`` / ``SelectiveScanMamba``: the production autograd Functions of
  R2GenCSR/VMamba/classification/models/vmamba.py:294-312 / 273-291 / 250-270 (same argument lists; backward
  returns the same 11-tuple).
* ``selective_scan_fn`` / ``SelectiveScanFn``: mamba_ssm's public op as the ARM mixers call it
  (CXPMRG_Bench_MambaXray_VL/arm/Finetuning/mamba_simple.py:693-704; wrapper semantics of
  R2GenCSR/VMamba/kernels/selective_scan/test_selective_scan.py:18-165).
* ``causal_conv1d_fn``, ``mamba_inner_fn``, ``mamba_inner_fn_no_out_proj``, ``bimamba_inner_fn``: the fused
  inner blocks of the (un-vendored) Vim/ARM fork of mamba_ssm, composed here from the conv / projection GEMMs and
  the CUDA scan exactly as the reference's own slow path states them (mamba_simple.py:665-709).  The scan is this
  repository's kernel; conv and GEMMs are library calls (next row of SURVEY.md section 8f).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import selective_scan_cuda, selective_scan_cuda_core, selective_scan_cuda_oflex

_fwd = torch.amp.custom_fwd(device_type="cuda")
_bwd = torch.amp.custom_bwd(device_type="cuda")


class SelectiveScanOflex(torch.autograd.Function):
    """vmamba.py:294-312."""

    @staticmethod
    @_fwd
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        ctx.delta_softplus = delta_softplus
        out, x, *rest = selective_scan_cuda_oflex.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1, oflex)
        ctx.has_hblk = len(rest) > 0                       # block states for the column-walk backward (scan_bwd_cw.cuh)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x, *rest[:1])
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x, *hb = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda_oflex.bwd(
            u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1, hblk=(hb[0] if hb else None))
        return (du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None)


class SelectiveScanCore(torch.autograd.Function):
    """vmamba.py:273-291."""

    @staticmethod
    @_fwd
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        ctx.delta_softplus = delta_softplus
        out, x, *rest = selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, 1)
        ctx.has_hblk = len(rest) > 0                       # block states for the column-walk backward (scan_bwd_cw.cuh)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x, *rest[:1])
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x, *hb = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda_core.bwd(
            u, delta, A, B, C, D, delta_bias, dout, x, ctx.delta_softplus, 1, hblk=(hb[0] if hb else None))
        return (du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None)


class SelectiveScanMamba(torch.autograd.Function):
    """vmamba.py:250-270 (mamba_ssm's module without the z gate)."""

    @staticmethod
    @_fwd
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        ctx.delta_softplus = delta_softplus
        out, x, *rest = selective_scan_cuda.fwd(u, delta, A, B, C, D, None, delta_bias, delta_softplus)
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
        return out

    @staticmethod
    @_bwd
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda.bwd(
            u, delta, A, B, C, D, None, delta_bias, dout, x, None, None, ctx.delta_softplus, False)
        return (du, ddelta, dA, dB, dC, dD, ddelta_bias, None, None, None, None)


class SelectiveScanFn(torch.autograd.Function):
    """mamba_ssm.ops.selective_scan_interface.SelectiveScanFn (semantics: test_selective_scan.py:20-147)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
        if u.stride(-1) != 1:
            u = u.contiguous()
        if delta.stride(-1) != 1:
            delta = delta.contiguous()
        if D is not None:
            D = D.contiguous()
        if B.stride(-1) != 1:
            B = B.contiguous()
        if C.stride(-1) != 1:
            C = C.contiguous()
        if z is not None and z.stride(-1) != 1:
            z = z.contiguous()
        ctx.squeeze_B = ctx.squeeze_C = False
        if B.dim() == 3:
            B = B.unsqueeze(1)
            ctx.squeeze_B = True
        if C.dim() == 3:
            C = C.unsqueeze(1)
            ctx.squeeze_C = True
        ctx.d_dtype = ctx.bias_dtype = None
        if D is not None and D.dtype != torch.float32:
            ctx.d_dtype = D.dtype
            D = D.float()
        if delta_bias is not None and delta_bias.dtype != torch.float32:
            ctx.bias_dtype = delta_bias.dtype
            delta_bias = delta_bias.float()
        out, x, *rest = selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)
        ctx.delta_softplus = delta_softplus
        ctx.has_z = z is not None
        last_state = x[:, :, -1, 1::2]  # (batch, dim, dstate)
        if not ctx.has_z:
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            return out if not return_last_state else (out, last_state)
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, out)
        out_z = rest[0]
        return out_z if not return_last_state else (out_z, last_state)

    @staticmethod
    def backward(ctx, dout, *args):
        if not ctx.has_z:
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            z = out = None
        else:
            u, delta, A, B, C, D, z, delta_bias, x, out = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda.bwd(
            u, delta, A, B, C, D, z, delta_bias, dout, x, out, None, ctx.delta_softplus, False)
        dz = rest[0] if ctx.has_z else None
        if ctx.squeeze_B:
            dB = dB.squeeze(1)
        if ctx.squeeze_C:
            dC = dC.squeeze(1)
        if dD is not None and ctx.d_dtype is not None:
            dD = dD.to(ctx.d_dtype)
        if ddelta_bias is not None and ctx.bias_dtype is not None:
            ddelta_bias = ddelta_bias.to(ctx.bias_dtype)
        return (du, ddelta, dA, dB, dC, dD, dz, ddelta_bias, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
    """mamba_ssm's op.  If return_last_state, returns (out, last_state (batch, dim, dstate)); the gradient of the
    last state is not propagated (as upstream)."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


# ---------------------------------------------------------------------------------------------------------
# Fused inner blocks of the Vim / ARM fork of mamba_ssm, composed (mamba_simple.py:665-709 is their definition).
_CONV_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _conv_rows(t):
    """(batch, dim, L) view with contiguous channel rows and a batch stride that is a multiple of 4 elements (the first
    half of `xz` qualifies as it is); anything else is made contiguous."""
    if t.stride(2) != 1 or t.stride(1) != t.shape[2] or t.stride(0) % 4 or t.data_ptr() % (4 * t.element_size()):
        t = t.contiguous()
    return t


class CausalConv1dFn(torch.autograd.Function):
    """Depth-wise causal conv1d (+ bias, + SiLU) through the C ABI (csrc/causal_conv1d.cu)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, silu):
        from . import _lib
        if x.dim() != 3 or weight.dim() != 2 or weight.shape[0] != x.shape[1]:
            raise RuntimeError("causal_conv1d_fn: x must be (batch, dim, seqlen) and weight (dim, width)")
        if not x.is_cuda:
            raise RuntimeError("causal_conv1d_fn: the B200 build has no CPU path (x must be a CUDA tensor)")
        x = _conv_rows(x)
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty(x.shape, device=x.device, dtype=x.dtype)
        B, D, L = x.shape
        rc = _lib.lib().mia_causal_conv1d_fwd(x.data_ptr(), w32.data_ptr(), 0 if b32 is None else b32.data_ptr(), y.data_ptr(), B, D, L,
                                              weight.shape[1], int(silu), _CONV_DT[x.dtype], x.stride(0), x.stride(1), y.stride(0),
                                              y.stride(1), torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"causal_conv1d_fwd: {_lib.lib().mia_conv_last_error().decode()} (code {rc})")
        ctx.save_for_backward(x, w32, b32 if b32 is not None else w32.new_empty(0))
        ctx.silu, ctx.has_bias, ctx.wdtype, ctx.bdtype = bool(silu), bias is not None, weight.dtype, None if bias is None else bias.dtype
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        from . import _lib
        x, w32, b32 = ctx.saved_tensors
        dy = _conv_rows(dy.to(x.dtype))
        B, D, L = x.shape
        dx = torch.empty(x.shape, device=x.device, dtype=x.dtype)
        dw = torch.empty_like(w32)
        db = torch.empty(D, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        rc = _lib.lib().mia_causal_conv1d_bwd(x.data_ptr(), w32.data_ptr(), b32.data_ptr() if ctx.has_bias else 0, dy.data_ptr(), dx.data_ptr(),
                                              dw.data_ptr(), 0 if db is None else db.data_ptr(), B, D, L, w32.shape[1], int(ctx.silu),
                                              _CONV_DT[x.dtype], x.stride(0), x.stride(1), dy.stride(0), dy.stride(1), dx.stride(0),
                                              dx.stride(1), torch.cuda.current_stream(x.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"causal_conv1d_bwd: {_lib.lib().mia_conv_last_error().decode()} (code {rc})")
        return dx, dw.to(ctx.wdtype), (db.to(ctx.bdtype) if db is not None else None), None


def causal_conv1d_fn(x, weight, bias=None, activation=None):
    """causal_conv1d.causal_conv1d_fn: depthwise causal conv, x (b, d, l), weight (d, w); mamba_simple.py:673-681."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    L = x.shape[-1]
    if L % 4:   # the kernels move 4-token pieces: pad on the right (a causal filter never looks there) and cut again
        return CausalConv1dFn.apply(F.pad(x, (0, 4 - L % 4)), weight, bias, activation is not None)[..., :L]
    return CausalConv1dFn.apply(x, weight, bias, activation is not None)


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
    """causal_conv1d.causal_conv1d_update (single-token decode step; only reached from ``Mamba.step``,
    mamba_simple.py:732-738, which no train.py of the reference calls): x (batch, dim), conv_state (batch, dim, width)
    rolled in place, weight (dim, width).  Not a hot path: stated with torch ops."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x
    out = torch.sum(conv_state * weight.to(conv_state.dtype), dim=-1)
    if bias is not None:
        out = out + bias.to(out.dtype)
    return (F.silu(out) if activation is not None else out).to(x.dtype)


def _inner_projections(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, B, C, B_proj_bias, C_proj_bias, d_state):
    if B is not None or C is not None:
        raise NotImplementedError("only input-dependent B / C (B=None, C=None), as at every call site of the reference")
    L = xz.shape[-1]
    x, z = xz.chunk(2, dim=1)
    cw = conv1d_weight.squeeze(1) if conv1d_weight.dim() == 3 else conv1d_weight
    x = causal_conv1d_fn(x, cw, conv1d_bias, "silu")
    from . import gemm as _gemm
    x_dbl = _gemm.linear(x.transpose(1, 2).reshape(-1, x.shape[1]), x_proj_weight)     # (b l, R + 2N): tcgen05 GEMM for bf16 / fp16
    R = delta_proj_weight.shape[1]
    N = (x_dbl.shape[1] - R) // 2 if d_state is None else d_state
    delta = (delta_proj_weight @ x_dbl[:, :R].t()).view(delta_proj_weight.shape[0], -1, L).transpose(0, 1)   # (b, d, l)
    Bm = x_dbl[:, R:R + N]
    Cm = x_dbl[:, R + N:R + 2 * N]
    if B_proj_bias is not None:
        Bm = Bm + B_proj_bias.to(Bm.dtype)
    if C_proj_bias is not None:
        Cm = Cm + C_proj_bias.to(Cm.dtype)
    Bm = Bm.view(-1, L, N).transpose(1, 2).unsqueeze(1).contiguous()                  # (b, 1, N, l)
    Cm = Cm.view(-1, L, N).transpose(1, 2).unsqueeze(1).contiguous()
    return x.contiguous(), z.contiguous(), delta.contiguous(), Bm, Cm


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None, D=None,
                               delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """(b, 2 d_inner, l) -> gated scan output (b, d_inner, l); call sites mamba_simple.py:450-511."""
    x, z, delta, Bm, Cm = _inner_projections(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, B, C,
                                             B_proj_bias, C_proj_bias, A.shape[1])
    return selective_scan_fn(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A,
                   B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """mamba_simple.py:650-663: the block above followed by out_proj; returns (b, l, d_model)."""
    y = mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D,
                                   delta_bias, B_proj_bias, C_proj_bias, delta_softplus)
    from . import gemm as _gemm
    return _gemm.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)


def bimamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A, A_b,
                     B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    """Vim's bidirectional block (mamba_simple.py:431-446): one set of projections, a forward scan with A and a
    scan over the flipped sequence with A_b, summed before out_proj."""
    x, z, delta, Bm, Cm = _inner_projections(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, B, C,
                                             B_proj_bias, C_proj_bias, A.shape[1])
    y_f = selective_scan_fn(x, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)
    fl = lambda t: t.flip([-1]).contiguous()
    y_b = selective_scan_fn(fl(x), fl(delta), A_b, fl(Bm), fl(Cm), D, z=fl(z), delta_bias=delta_bias, delta_softplus=delta_softplus)
    return F.linear((y_f + y_b.flip([-1])).transpose(1, 2), out_proj_weight, out_proj_bias)
