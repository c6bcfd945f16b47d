"""Helpers to read the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_DT = {"f32": torch.float32, "bf16": torch.bfloat16}


def scan_cases():
    z = np.load(os.path.join(GOLDEN, "scan_cases.npz"))
    out = []
    for spec in z["specs"]:
        tag, N, G, L, batch, dim, ddim, has_D, has_z, has_bias, softplus, dt = str(spec).split("|")
        dtype = _DT[dt]
        inp = {}
        for name in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "dout"):
            key = f"{tag}.in.{name}"
            if key in z.files:
                t = torch.from_numpy(z[key])
                inp[name] = t.to(dtype) if name in ("u", "delta", "B", "C", "z", "dout") else t
            else:
                inp[name] = None
        ref = {k.split(".ref.")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}.ref.")}
        out.append(dict(tag=tag, N=int(N), G=int(G), L=int(L), batch=int(batch), dim=int(dim), ddim=int(ddim),
                        softplus=bool(int(softplus)), dtype=dtype, inp=inp, ref=ref))
    return out


def c1_case():
    """BASELINE.json configs[0]: B=2 L=196 D=192 d_state=16 fp32 (inputs regenerated from seed 0)."""
    from tests.golden.make_golden import make_inputs
    z = np.load(os.path.join(GOLDEN, "scan_c1.npz"))
    inp = make_inputs(0, 2, 192, 196, 16, 1, 192, True, False, True, torch.float32)
    chk = float(sum(t.double().sum() for t in inp.values() if t is not None))
    assert abs(chk - float(z["in.checksum"][0])) < 1e-6 * max(1.0, abs(chk)), "torch RNG drifted: regenerate golden"
    ref = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ref.")}
    return inp, ref


def cross_scan():
    z = np.load(os.path.join(GOLDEN, "cross_scan.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}
