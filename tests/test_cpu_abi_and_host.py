"""CPU-only checks: the C-ABI library loads and exports what include/*.h declares, the ctypes mirror matches the C
struct, argument validation fails the way the reference's TORCH_CHECKs do, the drop-in registers the module names."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mia_selective_scan.h")


def _lib():
    from medical_image_analysis_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol():
    """Every function any include/*.h declares is exported by libmia_scan.so."""
    import glob
    lib = _lib()
    names = set()
    for hdr in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = open(hdr).read()
        body = text[text.index('extern "C"'):]
        names |= set(re.findall(r"\b(mia_[a-z_0-9]+)\s*\(", body))
    names = sorted(names)
    assert {"mia_gemm_tn", "mia_gemm_last_error", "mia_cs_last_error"} <= set(names)
    assert {"mia_selective_scan_fwd", "mia_selective_scan_bwd", "mia_selective_scan_bwd_workspace", "mia_last_error",
            "mia_abi_version", "mia_ss_num_chunks", "mia_ss_chunk_len", "mia_launch_count"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"libmia_scan.so does not export {n}"


def test_struct_layout_matches_header(tmp_path):
    from medical_image_analysis_b200._lib import MiaSSParams
    src = tmp_path / "sz.c"
    fields = [f for f, _ in MiaSSParams._fields_]
    lines = "\n".join(f'printf("{f} %zu\\n", offsetof(mia_ss_params, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mia_selective_scan.h"\nint main(){printf("size %zu\\n", '
                   'sizeof(mia_ss_params));' + lines + "return 0;}")
    exe = tmp_path / "sz"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["size"]) == ctypes.sizeof(MiaSSParams)
    for f in fields:
        assert int(out[f]) == getattr(MiaSSParams, f).offset, f


def test_chunk_geometry():
    lib = _lib()
    assert lib.mia_abi_version() == 2
    for L, (ch, n) in {1: (32, 1), 49: (64, 1), 128: (128, 1), 129: (256, 1), 196: (256, 1), 197: (256, 1), 256: (256, 1),
                       257: (256, 2), 6400: (256, 25)}.items():
        assert lib.mia_ss_chunk_len(L) == ch and lib.mia_ss_num_chunks(L) == n, L


def test_block_state_sizing_without_gpu():
    """ABI 2: mia_ss_block_state_floats is a pure function of the sizes (4 bytes per 16 row-tokens, d_state 1 with 32-row
    groups only); mia_ss_fwd_writes_block_states needs the device (and the buffers) and answers 0 without them."""
    from medical_image_analysis_b200._lib import MiaSSParams
    lib = _lib()
    p = MiaSSParams()
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups, p.delta_dim = 3, 128, 196, 1, 2, 128
    p.itype = p.otype = 2
    p.n_chunks = 1
    assert lib.mia_ss_block_state_floats(ctypes.byref(p)) == 3 * 128 * 13          # ceil(196 / 16) groups per row
    p.seqlen = 6400
    assert lib.mia_ss_block_state_floats(ctypes.byref(p)) == 3 * 128 * 400
    p.dstate = 16
    assert lib.mia_ss_block_state_floats(ctypes.byref(p)) == 0                      # the d_state 16 kernels do not use them
    p.dstate, p.dim, p.delta_dim, p.n_groups = 1, 96, 96, 2                         # 48 rows per group: not a multiple of 32
    assert lib.mia_ss_block_state_floats(ctypes.byref(p)) == 0
    assert lib.mia_ss_fwd_writes_block_states(ctypes.byref(p)) == 0
    assert lib.mia_ss_fwd_writes_block_states(None) == 0


def test_c_abi_validation_without_gpu():
    """Shape / dtype checks run before any CUDA call, so they are observable on a CPU box."""
    from medical_image_analysis_b200._lib import MiaSSParams
    lib = _lib()
    assert lib.mia_selective_scan_fwd(None, None) == -1
    p = MiaSSParams()
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups, p.delta_dim = 2, 6, 16, 1, 4, 6     # 6 % 4 != 0
    p.itype = p.otype = 0
    p.n_chunks = 1
    assert lib.mia_selective_scan_fwd(ctypes.byref(p), None) == -1
    assert b"n_groups" in lib.mia_last_error()
    p.n_groups, p.dstate = 2, 300
    assert lib.mia_selective_scan_fwd(ctypes.byref(p), None) == -1
    assert b"state dimension <= 256" in lib.mia_last_error()
    p.dstate, p.itype = 4, 7
    assert lib.mia_selective_scan_bwd(ctypes.byref(p), None) == -1
    assert b"float32, float16 or bfloat16" in lib.mia_last_error()


def test_gemm_validation_without_gpu():
    lib = _lib()
    buf = ctypes.create_string_buffer(64)
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.mia_gemm_tn(None, None, None, None, 4, 4, 8, 8, 8, 4, 2, 2, 0, None) == -1
    assert lib.mia_gemm_tn(ptr, ptr, None, ptr, 4, 4, 8, 8, 8, 4, 0, 0, 0, None) == -1           # fp32 inputs
    assert b"bf16 or fp16" in lib.mia_gemm_last_error()
    assert lib.mia_gemm_tn(ptr, ptr, None, ptr, 4, 4, 12, 12, 12, 4, 2, 2, 0, None) == -1        # pitch not a multiple of 8
    assert b"multiples of 8" in lib.mia_gemm_last_error()
    assert lib.mia_gemm_tn(ptr, ptr, None, ptr, 4, 4, 8, 8, 8, 4, 2, 2, 9, None) == -1
    assert b"activation" in lib.mia_gemm_last_error()


def test_host_checks_raise_runtime_error():
    from medical_image_analysis_b200 import scan_fwd
    u = torch.randn(1, 4, 8)
    A = -torch.rand(4, 2)
    B = torch.randn(1, 1, 2, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        scan_fwd(u, u.clone(), A, B, B.clone())
    with pytest.raises(RuntimeError, match="float32"):
        scan_fwd(u.double(), u.clone(), A, B, B.clone())


def test_no_silent_fallback_when_library_missing(monkeypatch, tmp_path):
    from medical_image_analysis_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "medical_image_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "ss_ref" not in text, f


def test_dropin_registers_reference_module_names():
    code = ("import medical_image_analysis_b200.dropin as d; d.install();"
            "import selective_scan_cuda_oflex as o, selective_scan_cuda_core as c, selective_scan_cuda as m;"
            "from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, mamba_inner_fn, bimamba_inner_fn, mamba_inner_fn_no_out_proj;"
            "from causal_conv1d import causal_conv1d_fn, causal_conv1d_update;"
            "from mamba_ssm.utils.generation import GenerationMixin; from mamba_ssm.utils.hf import load_config_hf, load_state_dict_hf;"
            "assert callable(o.fwd) and callable(o.bwd) and callable(c.fwd) and callable(m.bwd); print('ok')")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert out.strip().endswith("ok")


def test_causal_conv1d_has_no_cpu_path():
    """causal_conv1d_fn is a CUDA kernel behind the C ABI (parity vs mamba_simple.py:673 in tests/test_causal_conv1d_gpu.py);
    like the scan it must fail loudly, not fall back, when it cannot run on the GPU."""
    import pytest
    from medical_image_analysis_b200.selective_scan_interface import causal_conv1d_fn
    conv = torch.nn.Conv1d(6, 6, 4, groups=6, padding=3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        causal_conv1d_fn(torch.randn(2, 6, 20), conv.weight.squeeze(1), conv.bias, "silu")


def test_dwconv2d_has_no_cpu_path():
    """SS2D's depth-wise conv kernel (csrc/dwconv2d.cu) is CUDA-only as well: a CPU tensor must raise, not fall back."""
    import pytest
    from medical_image_analysis_b200.vmamba import DWConv2dFn
    conv = torch.nn.Conv2d(4, 4, 3, padding=1, groups=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        DWConv2dFn.apply(torch.randn(1, 4, 5, 5), conv.weight, conv.bias, True)
    with pytest.raises(RuntimeError, match="weight"):
        DWConv2dFn.apply(torch.randn(1, 4, 5, 5), torch.randn(4, 1, 5, 5), None, True)
