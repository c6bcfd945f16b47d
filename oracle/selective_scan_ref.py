"""Torch-CPU restatement of the reference's selective-scan oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: PINNED --
tests/test_oracle_golden.py checks every function here against vectors that
tests/golden/make_golden.py produced by executing the reference's own
``selective_scan_ref`` (extracted with ``ast`` from
/root/reference/R2GenCSR/VMamba/kernels/selective_scan/test_selective_scan.py:168-234,
the module itself cannot be imported because it imports the CUDA extensions at
lines 356-357) and torch autograd through it.

The functions follow the reference's arithmetic order (fp32 upcast, softplus
on delta+bias, exp(delta (x) A), delta*B*u, serial recurrence, <h, C>, D skip,
silu(z) gate, cast back) so that ``bench.py``'s CPU-baseline leg times the same
algorithm the reference would run on a CPU.  Real-valued A only: the CUDA op
the repo replaces rejects complex weights (selective_scan_oflex.cpp:158).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False):
    """Follows test_selective_scan.py:168-234.

    u, delta: (batch, dim, L); A: (dim, N) fp32; B, C: (batch, N, L) or
    (batch, G, N, L); D, delta_bias: (dim,) fp32; z: (batch, dim, L).
    Returns out (batch, dim, L) in u's dtype [, last_state (batch, dim, N) fp32].
    """
    in_dtype = u.dtype
    uf = u.float()
    dt = delta.float()
    if delta_bias is not None:                      # ref :195-196
        dt = dt + delta_bias.float().unsqueeze(-1)
    if delta_softplus:                              # ref :197-198
        dt = F.softplus(dt)
    batch, dim, L = uf.shape
    N = A.shape[1]
    Bf, Cf = B.float(), C.float()
    if Bf.dim() == 3:
        Bf = Bf.unsqueeze(1)
    if Cf.dim() == 3:
        Cf = Cf.unsqueeze(1)
    # groups -> per-channel (ref :210, :212-213)
    Bf = Bf.repeat_interleave(dim // Bf.shape[1], dim=1)    # (b, dim, N, L)
    Cf = Cf.repeat_interleave(dim // Cf.shape[1], dim=1)
    decay = torch.exp(dt.unsqueeze(-1) * A.float().view(1, dim, 1, N))       # (b, dim, L, N)  ref :203
    drive = (dt * uf).unsqueeze(-1) * Bf.permute(0, 1, 3, 2)                 # (b, dim, L, N)  ref :205-211
    h = A.new_zeros((batch, dim, N), dtype=torch.float32)
    ys = []
    for l in range(L):                                                       # ref :215-227
        h = decay[:, :, l] * h + drive[:, :, l]
        ys.append((h * Cf[:, :, :, l]).sum(-1))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + uf * D.float().view(1, dim, 1)             # ref :230
    if z is not None:                                                        # ref :231-232
        out = out * F.silu(z)       # NOT upcast: the reference evaluates silu in z's own dtype
    out = out.to(in_dtype)                                                   # ref :233
    return (out, h) if return_last_state else out


def expand_delta_groups(delta, delta_bias, dim):
    """oflex delta-groups: row d of the scan uses delta row d // (dim/delta_dim)
    (selective_scan_oflex.cpp:59, fwd kernel :93-97); this is how the reference
    test builds its comparison input (test_selective_scan.py:453-457)."""
    rep = dim // delta.shape[1]
    d_full = delta.repeat_interleave(rep, dim=1)
    b_full = None if delta_bias is None else delta_bias.repeat_interleave(rep, dim=0)
    return d_full, b_full


def fold_delta_group_grads(ddelta_full, ddelta_bias_full, delta_dim):
    """Adjoint of expand_delta_groups (test_selective_scan.py:510-517,
    selective_scan_oflex.cpp:348-353)."""
    b, dim, L = ddelta_full.shape
    rep = dim // delta_dim
    dd = ddelta_full.view(b, delta_dim, rep, L).sum(2)
    db = None if ddelta_bias_full is None else ddelta_bias_full.view(delta_dim, rep).sum(1)
    return dd, db


def selective_scan_ref_fwd_bwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout):
    """Forward + torch-autograd backward through the restatement, the way the
    reference test does it (test_selective_scan.py:483-485).  delta may have
    fewer rows than u (oflex delta groups).  Returns a dict of out/last_state
    and all gradients as fp32 CPU tensors."""
    dim = u.shape[1]
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_()
    u_, A_, B_, C_, D_, z_ = map(leaf, (u, A, B, C, D, z))
    d_small, b_small = leaf(delta), leaf(delta_bias)
    if delta.shape[1] != dim:
        d_full, b_full = expand_delta_groups(d_small, b_small, dim)
    else:
        d_full, b_full = d_small, b_small
    out, last = selective_scan_ref(u_, d_full, A_, B_, C_, D_, z_, b_full,
                                   delta_softplus, return_last_state=True)
    out.backward(dout.to(out.dtype))
    g = lambda t: None if t is None else t.grad
    return dict(out=out.detach(), last_state=last.detach(), du=g(u_), ddelta=g(d_small),
                dA=g(A_), dB=g(B_), dC=g(C_), dD=g(D_), dz=g(z_), ddelta_bias=g(b_small))


def selective_scan_ref_v2(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
    """Follows test_selective_scan.py:237-306: the low-precision behaviour model -- every operand is first cast to u's
    dtype (:253-259) and the whole recurrence runs in that dtype (real A, variable B / C of rank 3 or 4).  Not a parity
    target of the kernels (they compute in fp32 like selective_scan_ref); restated so that the reference's second oracle
    has a counterpart, pinned to vectors of the reference's own function (tests/golden/scan_ref_v2.npz)."""
    dt_in = u.dtype
    A, B, C = A.to(dt_in), B.to(dt_in), C.to(dt_in)
    D = None if D is None else D.to(dt_in)
    z = None if z is None else z.to(dt_in)
    delta = delta.to(dt_in)
    if delta_bias is not None:
        delta = delta + delta_bias.to(dt_in)[..., None]
    if delta_softplus:
        delta = F.softplus(delta)
    batch, dim, N = u.shape[0], A.shape[0], A.shape[1]
    dA = torch.exp(torch.einsum("bdl,dn->bdln", delta, A))
    if B.dim() == 3:
        dBu = torch.einsum("bdl,bnl,bdl->bdln", delta, B, u)
    else:
        dBu = torch.einsum("bdl,bdnl,bdl->bdln", delta, B.repeat_interleave(dim // B.shape[1], dim=1), u)
    if C.dim() == 4:
        C = C.repeat_interleave(dim // C.shape[1], dim=1)
    h = A.new_zeros((batch, dim, N))
    ys = []
    for i in range(u.shape[2]):
        h = dA[:, :, i] * h + dBu[:, :, i]
        ys.append(torch.einsum("bdn,bn->bd", h, C[:, :, i]) if C.dim() == 3 else torch.einsum("bdn,bdn->bd", h, C[:, :, :, i]))
    out = torch.stack(ys, dim=2)
    if D is not None:
        out = out + u * D.unsqueeze(-1)
    if z is not None:
        out = out * F.silu(z)
    out = out.to(dt_in)
    return (out, h.float()) if return_last_state else out
