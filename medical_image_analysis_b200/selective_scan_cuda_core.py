"""Drop-in for ``selective_scan_cuda_core`` (.../cus/selective_scan.cpp:157-164, 241-250, 351-354): the oflex
op with the output in the input dtype and no ``out_float`` argument."""
from __future__ import annotations

from .scan_op import scan_bwd, scan_fwd


def fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1):
    out, x, _, hblk = scan_fwd(u, delta, A, B, C, D, None, delta_bias, delta_softplus, False, want_block_states=True)
    # a third element rides in the reference's `out, x, *rest = ...fwd(...)` (vmamba.py:299): block states for `bwd(..., hblk=)`
    return [out, x] if hblk is None else [out, x, hblk]


def bwd(u, delta, A, B, C, D=None, delta_bias=None, dout=None, x=None, delta_softplus=False, nrows=1, hblk=None):
    du, dd, dA, dB, dC, dD, dbias, _ = scan_bwd(u, delta, A, B, C, D, None, delta_bias, dout, x, None, delta_softplus, hblk=hblk)
    return [du, dd, dA, dB, dC, dD, dbias]
