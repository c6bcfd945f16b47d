"""ctypes front-end of oracle/ss_ref.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libss_ref.so")
    src = os.path.join(_HERE, "ss_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libss_ref.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f32(t):
    return None if t is None else np.ascontiguousarray(t.detach().float().cpu().numpy())


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _norm_bc(B):
    return B.unsqueeze(1) if B.dim() == 3 else B


def fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    """Returns (out, out_z or None, last_state) as fp32 CPU tensors (out is the pre-gate y + D u)."""
    B, C = _norm_bc(B), _norm_bc(C)
    batch, dim, L = u.shape
    N, G, ddim = A.shape[1], B.shape[1], delta.shape[1]
    a = [_f32(t) for t in (u, delta, A, B, C, D, z, delta_bias)]
    out = np.empty((batch, dim, L), np.float32)
    out_z = np.empty_like(out) if z is not None else None
    last = np.empty((batch, dim, N), np.float32)
    _lib().ss_ref_fwd(*[_p(x) for x in a], ctypes.c_int(int(delta_softplus)), batch, dim, L, N, G, ddim,
                      _p(out), _p(out_z), _p(last))
    t = torch.from_numpy
    return t(out), (t(out_z) if out_z is not None else None), t(last)


def bwd(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus=False):
    """Returns dict of fp32 CPU gradients (ddelta / ddelta_bias already folded over delta groups)."""
    B, C = _norm_bc(B), _norm_bc(C)
    batch, dim, L = u.shape
    N, G, ddim = A.shape[1], B.shape[1], delta.shape[1]
    a = [_f32(t) for t in (u, delta, A, B, C, D, z, delta_bias, dout)]
    du = np.empty((batch, dim, L), np.float32)
    dd = np.empty((batch, ddim, L), np.float32)
    dA = np.empty((dim, N), np.float32)
    dB = np.empty((batch, G, N, L), np.float32)
    dC = np.empty_like(dB)
    dD = np.empty((dim,), np.float32) if D is not None else None
    db = np.empty((ddim,), np.float32) if delta_bias is not None else None
    dz = np.empty_like(du) if z is not None else None
    _lib().ss_ref_bwd(*[_p(x) for x in a], ctypes.c_int(int(delta_softplus)), batch, dim, L, N, G, ddim,
                      _p(du), _p(dd), _p(dA), _p(dB), _p(dC), _p(dD), _p(db), _p(dz))
    t = lambda x: None if x is None else torch.from_numpy(x)
    return dict(du=t(du), ddelta=t(dd), dA=t(dA), dB=t(dB), dC=t(dC), dD=t(dD), ddelta_bias=t(db), dz=t(dz))
