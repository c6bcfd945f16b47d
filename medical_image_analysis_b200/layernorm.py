"""``mamba_ssm.ops.triton.layernorm`` surface (``RMSNorm``, ``rms_norm_fn``, ``layer_norm_fn``) for the drop-in.

The ARM / Vim sub-projects import these three names (arm/Finetuning/models_mamba.py:24, mamba_simple.py:30, pretrain/
models_pretrain.py:26).  Every registered ARM factory passes ``rms_norm=True`` (models_mamba.py:401, 415, 428), so
``create_block`` builds ``partial(RMSNorm, eps=...)`` (:152-154) -- the class must exist for the models to be
constructed, even though the ``Block`` that is actually instantiated (:86-116) normalises with ``nn.LayerNorm``.  The
functional forms are only reached from the unused ``mamba_simple.Block`` (:807-862) when ``fused_add_norm`` is set.

Upstream these are Triton kernels; the north star forbids Triton and none of them is on the measured hot path, so
they are stated with torch ops (fp32 statistics, one rounding), with the upstream signatures and return conventions
(``prenorm`` -> ``(y, residual_out)``, ``residual_in_fp32``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm):
    if residual is None:
        res = x.float() if residual_in_fp32 else x
    elif residual_in_fp32:
        res = x.float() + residual.float()
    else:
        res = (x.float() + residual.float()).to(x.dtype)
    xf = res.float()
    if is_rms_norm:
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        y = y * weight.float()
        if bias is not None:
            y = y + bias.float()
    else:
        y = F.layer_norm(xf, (xf.shape[-1],), weight.float(), None if bias is None else bias.float(), eps)
    y = y.to(x.dtype)
    return (y, res) if prenorm else y


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False, is_rms_norm=False):
    """Fused (residual add +) LayerNorm / RMSNorm.  Returns y, or (y, x + residual) when ``prenorm``."""
    return _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms_norm)


def rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
    return _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, True)


class RMSNorm(nn.Module):
    """``mamba_ssm.ops.triton.layernorm.RMSNorm``: weight-only RMS normalisation over the last dimension."""

    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm,
                           residual_in_fp32=residual_in_fp32)
