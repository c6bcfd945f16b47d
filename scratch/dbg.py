import torch, sys
sys.path.insert(0,'/root/repo')
from medical_image_analysis_b200 import scan_fwd, scan_bwd
from tests.golden_util import scan_cases
for case in scan_cases():
    i, ref = case["inp"], case["ref"]
    g = {k: (None if v is None else v.cuda()) for k, v in i.items()}
    out, x, out_z = scan_fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["z"], g["delta_bias"], case["softplus"], False)
    grads = scan_bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], g["z"], g["delta_bias"], g["dout"], x, out if g["z"] is not None else None, case["softplus"])
    names = ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz")
    msg=[]
    for name, got in zip(names, grads):
        if got is not None:
            e=(got.float().cpu()-ref[name]).abs().max().item(); s=ref[name].abs().max().item()
            if e > 3e-2*max(1,s): msg.append(f"{name}:{e:.3g}/{s:.3g}")
    print(case["tag"], "N",case["N"],"G",case["G"],"L",case["L"],"dd",case["ddim"],case["dtype"], msg)
