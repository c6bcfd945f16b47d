"""GPU-side debugging aid: the d_state 16 backward (scan_bwd_rowsn.cuh) against the C oracle, per-tensor worst errors.
Run under compute-sanitizer for the memory check:  compute-sanitizer --tool memcheck python tools/rowsn_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medical_image_analysis_b200 import scan_bwd, scan_fwd  # noqa: E402
from oracle import ss_ref_c  # noqa: E402


def case(batch, dim, L, G, has_z, dtype, seed=0):
    N = 16
    g = torch.Generator().manual_seed(seed)
    A = -0.5 * torch.rand(dim, N, generator=g)
    B = torch.randn(batch, G, N, L, generator=g).to(dtype)
    C = torch.randn(batch, G, N, L, generator=g).to(dtype)
    D = torch.randn(dim, generator=g)
    z = torch.randn(batch, dim, L, generator=g).to(dtype) if has_z else None
    bias = 0.5 * torch.rand(dim, generator=g)
    u = torch.randn(batch, dim, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(batch, dim, L, generator=g)).to(dtype)
    dout = torch.randn(batch, dim, L, generator=g).to(dtype)
    cu = lambda t: None if t is None else t.cuda()
    out, x, out_z = scan_fwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(z), cu(bias), True, False)
    grads = scan_bwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(z), cu(bias), cu(dout), x, out if has_z else None, True)
    torch.cuda.synchronize()
    ref = ss_ref_c.bwd(u, delta, A, B, C, D, z, bias, dout, True)
    names = ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz")
    line = []
    for n, gt in zip(names, grads):
        if gt is None:
            continue
        r = ref[n]
        err = (gt.float().cpu() - r).abs().max().item()
        line.append(f"{n}:{err:.2e}/{r.abs().max().item():.2e}")
    print(f"b{batch} d{dim} L{L} G{G} z{int(has_z)} {str(dtype)[6:]}: " + " ".join(line), flush=True)


if __name__ == "__main__":
    for dt in (torch.float32, torch.bfloat16):
        case(2, 64, 196, 2, False, dt)
        case(1, 32, 197, 1, True, dt)
        case(2, 64, 5, 1, False, dt)
        case(3, 96, 256, 3, True, dt)
        case(2, 768, 197, 1, True, dt)
