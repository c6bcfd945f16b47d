"""Drop-in for the reference's pybind module ``selective_scan_cuda_oflex``
(R2GenCSR/VMamba/kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp:357-360):
same two entry points, same positional arguments, same result lists, RuntimeError on a failed check."""
from __future__ import annotations

from .scan_op import scan_bwd, scan_fwd


def fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, out_float=True):
    """selective_scan_fwd (selective_scan_oflex.cpp:143-231) -> [out, x].  `nrows` is accepted and ignored
    (the reference only instantiates nrows == 1, :223-227)."""
    out, x, _, hblk = scan_fwd(u, delta, A, B, C, D, None, delta_bias, delta_softplus, bool(out_float), want_block_states=True)
    # a third element rides in the reference's `out, x, *rest = ...fwd(...)` (vmamba.py:299): block states for `bwd(..., hblk=)`
    return [out, x] if hblk is None else [out, x, hblk]


def bwd(u, delta, A, B, C, D=None, delta_bias=None, dout=None, x=None, delta_softplus=False, nrows=1, hblk=None):
    """selective_scan_bwd (selective_scan_oflex.cpp:233-355) -> [du, ddelta, dA, dB, dC, dD, ddelta_bias];
    ddelta / ddelta_bias are already folded over the delta group (:348-353), dB / dC are in the input dtype (:347)."""
    du, dd, dA, dB, dC, dD, dbias, _ = scan_bwd(u, delta, A, B, C, D, None, delta_bias, dout, x, None, delta_softplus, hblk=hblk)
    return [du, dd, dA, dB, dC, dD, dbias]
