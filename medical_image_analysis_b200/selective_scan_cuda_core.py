"""Drop-in for ``selective_scan_cuda_core`` (.../cus/selective_scan.cpp:157-164, 241-250, 351-354): the oflex
op with the output in the input dtype and no ``out_float`` argument."""
from __future__ import annotations

from .scan_op import scan_bwd, scan_fwd


def fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1):
    out, x, _ = scan_fwd(u, delta, A, B, C, D, None, delta_bias, delta_softplus, False)
    return [out, x]


def bwd(u, delta, A, B, C, D=None, delta_bias=None, dout=None, x=None, delta_softplus=False, nrows=1):
    du, dd, dA, dB, dC, dD, dbias, _ = scan_bwd(u, delta, A, B, C, D, None, delta_bias, dout, x, None, delta_softplus)
    return [du, dd, dA, dB, dC, dD, dbias]
