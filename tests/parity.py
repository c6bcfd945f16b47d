"""Parity criteria shared by the GPU tests (north star: 1e-3 rtol bf16, 1e-5 fp32, against the reference's own path).

Two criteria, both element-wise:

* ``cmp_stored``  -- a tensor the kernel ROUNDS to a storage dtype (bf16 / fp16 outputs and gradients).  The fp64 oracle is
  rounded to that dtype first ("identical single rounding"), then the CUDA value must sit within ``n_ulp`` units
  in the last place of the storage dtype of it, plus an absolute term tied to the RMS of the reference tensor (NOT to its
  maximum): ``atol = atol_rms * RMS(ref)``.  One bf16 ulp is 2^-7 relative: the north star's 1e-3 cannot be met by ANY
  implementation that stores bf16 (half an ulp of rounding alone is 2e-3), so the bar is "the same rounding as the
  reference would produce, give or take a rounding boundary", which is what 1 ulp of the rounded oracle states.
* ``cmp_f32``     -- fp32 tensors (fp32 inputs, the fp32 "oflex" outputs of low-precision inputs, weight gradients):
  ``|got - ref| <= rtol * |ref| + atol_rms * RMS(ref)`` with rtol = 1e-5.

Setting MIA_PARITY_LOG=<file> appends one JSON line per comparison with the worst ratio error / tolerance (calibration
evidence for the numbers above; written by the GPU runs under gpurun_out/).
"""
import json
import os

import torch

_MANT = {torch.bfloat16: 8, torch.float16: 11, torch.float32: 24}      # significand bits incl. the hidden one
_MIN_EXP = {torch.bfloat16: -126, torch.float16: -14, torch.float32: -126}


def ulp_of(ref: torch.Tensor, dtype) -> torch.Tensor:
    """Spacing of ``dtype`` at the magnitude of each element of ``ref`` (float64 tensor)."""
    a = ref.abs().double()
    e = torch.floor(torch.log2(torch.clamp(a, min=2.0 ** _MIN_EXP[dtype])))
    return torch.pow(2.0, e - (_MANT[dtype] - 1))


def _rms(t: torch.Tensor) -> float:
    return float(t.double().pow(2).mean().sqrt()) if t.numel() else 0.0


def _report(what, err, tol, ref):
    ratio = err / tol
    worst = float(ratio.max()) if ratio.numel() else 0.0
    path = os.environ.get("MIA_PARITY_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps({"what": what, "worst_ratio": worst, "max_err": float(err.max()) if err.numel() else 0.0,
                                "rms_ref": _rms(ref), "n": int(err.numel())}) + "\n")
    bad = ratio > 1.0
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance, worst err/tol {worst:.2f}, "
                                 f"max err {float(err.max()):.3e}, RMS(ref) {_rms(ref):.3e}")


def cmp_stored(got: torch.Tensor, ref: torch.Tensor, dtype, what: str, n_ulp: float = 1.0, atol_rms: float = 1e-3):
    """``got`` is stored in ``dtype``; ``ref`` is the fp64 / fp32 oracle value before any rounding."""
    assert got.dtype == dtype, (what, got.dtype, dtype)
    got64, ref64 = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got64.shape == ref64.shape, (what, got64.shape, ref64.shape)
    assert torch.isfinite(got64).all(), f"{what}: non-finite values"
    refq = ref64.to(dtype).double()                              # the oracle after the same single rounding
    err = (got64 - refq).abs()
    tol = n_ulp * ulp_of(refq, dtype) + atol_rms * _rms(ref64)
    _report(what, err, tol, ref64)


def cmp_f32(got: torch.Tensor, ref: torch.Tensor, what: str, rtol: float = 1e-5, atol_rms: float = 2e-5):
    got64, ref64 = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got64.shape == ref64.shape, (what, got64.shape, ref64.shape)
    assert torch.isfinite(got64).all(), f"{what}: non-finite values"
    err = (got64 - ref64).abs()
    tol = rtol * ref64.abs() + atol_rms * _rms(ref64) + 1e-30
    _report(what, err, tol, ref64)


def long_row_scale(L: int) -> float:
    """fp32 criterion for rows longer than one 256-token chunk.  h_t is a product of up to L factors a_s = exp(dl_s A): a
    relative error eps per factor (fp32 rounding, ~1e-7) grows to ~n eps over the n <= L tokens a slowly decaying row
    remembers, in ANY fp32 evaluation.  Measured here with the reference's own fp32 path (oracle/selective_scan_ref.py, exact
    libm exp) against the fp64 oracle, reference generators, 256 rows: worst err / (1e-5 |ref| + 2e-5 RMS) = 0.16 at
    L = 196, 0.52 at L = 2048, 2.24 at L = 6400.  The tolerance therefore scales with the number of chunks."""
    return max(1.0, L / 256.0)


def cmp_auto(got: torch.Tensor, ref: torch.Tensor, what: str, n_ulp: float = 1.0, f32_scale: float = 1.0):
    """fp32 tensors by ``cmp_f32`` (tolerances times ``f32_scale``, see long_row_scale), bf16 / fp16 tensors by
    ``cmp_stored`` (``n_ulp`` units in the last place).  Stored tensors of long rows: the fp32 value IN FRONT of the rounding
    carries the noise of a recurrence over up to L tokens, which shows where terms ~RMS cancel to a small element (measured,
    gpurun r2v, L = 6400: 1 of 19.66 M ddelta elements 3 ulp = 3.3e-3 RMS off at |value| = 0.2 RMS) -> the absolute term grows
    with the chunk count too, a fifth as fast as the fp32 criterion."""
    if got.dtype == torch.float32:
        return cmp_f32(got, ref, what, rtol=1e-5 * f32_scale, atol_rms=2e-5 * f32_scale)
    return cmp_stored(got, ref, got.dtype, what, n_ulp=n_ulp, atol_rms=1e-3 * max(1.0, f32_scale / 5.0))
