"""B200-native (sm_100a) selective scan for the Vision-Mamba models of Event-AHU/Medical_Image_Analysis.

Only the hot path lives here: the CUDA kernels + C ABI (``csrc/``, ``libmia_scan.so``) and the host-side
mirrors of the reference's operator surface (``selective_scan_cuda_oflex`` / ``_core`` / ``selective_scan_cuda``
modules, the ``SelectiveScan*`` autograd Functions, ``selective_scan_fn``).
"""
from . import _lib  # noqa: F401
from .scan_op import num_chunks, scan_bwd, scan_fwd  # noqa: F401

__all__ = ["scan_fwd", "scan_bwd", "num_chunks"]
