"""Make the reference's sub-projects import this library under the module names they expect.

    import medical_image_analysis_b200.dropin as dropin; dropin.install()

registers ``selective_scan_cuda_oflex``, ``selective_scan_cuda_core``, ``selective_scan_cuda`` (imported by
R2GenCSR/VMamba/classification/models/vmamba.py:133-155 and the kernel tests), and a minimal ``mamba_ssm`` /
``causal_conv1d`` package tree (imported by */arm/Finetuning/mamba_simple.py:14-32, models_mamba.py:19-24 and
*/pretrain/*.py) in ``sys.modules`` -- no files are written and nothing is monkey-patched inside the reference.
"""
from __future__ import annotations

import sys
import types

from . import layernorm, selective_scan_cuda, selective_scan_cuda_core, selective_scan_cuda_oflex, selective_scan_interface as ssi


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so that sub-module imports resolve through sys.modules
    sys.modules[name] = m
    return m


def install(force: bool = False) -> None:
    for name, mod in (("selective_scan_cuda_oflex", selective_scan_cuda_oflex),
                      ("selective_scan_cuda_core", selective_scan_cuda_core),
                      ("selective_scan_cuda", selective_scan_cuda)):
        if force or name not in sys.modules:
            sys.modules[name] = mod
    if not force and "mamba_ssm" in sys.modules:
        return
    iface = _module("mamba_ssm.ops.selective_scan_interface",
                    selective_scan_fn=ssi.selective_scan_fn, SelectiveScanFn=ssi.SelectiveScanFn,
                    mamba_inner_fn=ssi.mamba_inner_fn, mamba_inner_fn_no_out_proj=ssi.mamba_inner_fn_no_out_proj,
                    bimamba_inner_fn=ssi.bimamba_inner_fn)
    # models_mamba.py:24 / mamba_simple.py:30: every arm_*_pz16 factory passes rms_norm=True, so RMSNorm must be a class
    ln = _module("mamba_ssm.ops.triton.layernorm", RMSNorm=layernorm.RMSNorm, layer_norm_fn=layernorm.layer_norm_fn,
                 rms_norm_fn=layernorm.rms_norm_fn)
    tri = _module("mamba_ssm.ops.triton", layernorm=ln)     # selective_state_update (only Mamba.step) stays an ImportError
    ops = _module("mamba_ssm.ops", selective_scan_interface=iface, triton=tri)
    class GenerationMixin:  # imported (unused) by models_mamba.py:20 / models_pretrain.py:20
        pass

    def _no_hub(*_a, **_k):
        raise NotImplementedError("mamba_ssm.utils.hf is not part of the B200 drop-in (imported but unused by the reference)")

    gen = _module("mamba_ssm.utils.generation", GenerationMixin=GenerationMixin)
    hf = _module("mamba_ssm.utils.hf", load_config_hf=_no_hub, load_state_dict_hf=_no_hub)
    utils = _module("mamba_ssm.utils", generation=gen, hf=hf)
    root = _module("mamba_ssm", ops=ops, utils=utils, __version__="0+b200")
    cc = _module("causal_conv1d", causal_conv1d_fn=ssi.causal_conv1d_fn, causal_conv1d_update=ssi.causal_conv1d_update)
    del root, cc
