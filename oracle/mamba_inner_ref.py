"""CPU restatement of the ARM/Vim Mamba mixer's inner block (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The fused ops ``mamba_inner_fn`` / ``mamba_inner_fn_no_out_proj`` / ``bimamba_inner_fn`` live in an un-vendored fork
of mamba_ssm.  Parity status: UNPINNED against that fork (it is absent from /root/reference and the
reference holds no test vectors for it); what is restated here is the reference's OWN slow path
(CXPMRG_Bench_MambaXray_VL/arm/Finetuning/mamba_simple.py:665-709), which the reference treats as the definition of
the fused op, on top of the pinned scan oracle.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .selective_scan_ref import selective_scan_ref


def mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, D, delta_bias,
                    out_proj_weight=None, out_proj_bias=None):
    """mamba_simple.py:665-709 with conv_state = ssm_state = None.  xz: (b, 2 d_inner, l)."""
    L = xz.shape[-1]
    d_inner = xz.shape[1] // 2
    dt_rank = delta_proj_weight.shape[1]
    d_state = A.shape[1]
    x, z = xz.chunk(2, dim=1)                                                                       # :665
    x = F.silu(F.conv1d(x, conv1d_weight, conv1d_bias, padding=conv1d_weight.shape[-1] - 1, groups=d_inner)[..., :L])  # :673
    x_dbl = F.linear(x.transpose(1, 2).reshape(-1, d_inner), x_proj_weight)                         # :686
    dt, B, C = torch.split(x_dbl, [dt_rank, d_state, d_state], dim=-1)                              # :687
    dt = (delta_proj_weight @ dt.t()).view(d_inner, -1, L).transpose(0, 1)                          # :688-689
    B = B.view(-1, L, d_state).transpose(1, 2).contiguous()                                         # :690
    C = C.view(-1, L, d_state).transpose(1, 2).contiguous()                                         # :691
    y = selective_scan_ref(x, dt, A, B, C, D.float(), z=z, delta_bias=delta_bias.float(), delta_softplus=True)  # :693-704
    if out_proj_weight is None:
        return y
    return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)                              # :708-709
