"""CPU restatement of VMamba's CrossScan / CrossMerge / cross_selective_scan / SS2D.forwardv2
(TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

CrossScan / CrossMerge are PINNED to tests/golden/cross_scan.npz, produced by executing the reference's own classes
(R2GenCSR/VMamba/classification/models/vmamba.py:25-67).  cross_selective_scan / forwardv2 (vmamba.py:318-427, 1110-1129)
have no reference-held vectors (vmamba.py cannot be imported here: timm / fvcore / triton csm) -> parity of the module
composition is "unpinned"; it is restated from the source on top of the pinned pieces.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .selective_scan_ref import selective_scan_ref


def cross_scan_ref(x):
    """vmamba.py:28-35."""
    B, C, H, W = x.shape
    xs = x.new_empty((B, 4, C, H * W))
    xs[:, 0] = x.flatten(2, 3)
    xs[:, 1] = x.transpose(2, 3).flatten(2, 3)
    xs[:, 2:4] = torch.flip(xs[:, 0:2], dims=[-1])
    return xs


def cross_merge_ref(ys, H, W):
    """vmamba.py:51-57; ys (B, 4, D, L) -> (B, D, L)."""
    B, K, D, L = ys.shape
    ys = ys[:, 0:2] + ys[:, 2:4].flip(dims=[-1]).view(B, 2, D, -1)
    return ys[:, 0] + ys[:, 1].view(B, -1, W, H).transpose(2, 3).contiguous().view(B, D, -1)


def cross_selective_scan_ref(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, out_norm, to_bf16=True):
    """vmamba.py:318-427, einsum path, channel_last, out_norm_shape v0, delta_softplus=True."""
    B, D, H, W = x.shape
    N = A_logs.shape[1]
    K, _, R = dt_projs_weight.shape
    L = H * W
    xs = cross_scan_ref(x)                                                             # :385
    x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs, x_proj_weight)               # :386
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                                 # :389
    dts = torch.einsum("b k r l, k d r -> b k d l", dts, dt_projs_weight)              # :390
    As = -torch.exp(A_logs.float())                                                    # :394
    ys = selective_scan_ref(xs.reshape(B, -1, L).float(), dts.reshape(B, -1, L).float(), As, Bs.contiguous().float(),
                            Cs.contiguous().float(), Ds.float(), None, dt_projs_bias.reshape(-1).float(), True)
    y = cross_merge_ref(ys.view(B, K, -1, L), H, W)                                    # :410
    if to_bf16:
        y = y.to(torch.bfloat16).float()                                               # :420
    y = out_norm(y.transpose(1, 2).contiguous()).view(B, H, W, -1)                     # :423-425
    return y


def ss2d_forward_ref(m, x):
    """SS2D.forwardv2 (vmamba.py:1110-1129) on a module holding the reference's parameter names."""
    x = m.in_proj(x)
    z = None
    if not m.disable_z:
        x, z = x.chunk(2, dim=-1)
        if not m.disable_z_act:
            z = m.act(z)
    x = x.permute(0, 3, 1, 2).contiguous()
    if m.d_conv > 1:
        x = m.conv2d(x)
    x = m.act(x)
    y = cross_selective_scan_ref(x, m.x_proj_weight, m.dt_projs_weight, m.dt_projs_bias, m.A_logs, m.Ds, m.out_norm)
    if z is not None:
        y = y * z
    return m.dropout(m.out_proj(y))
