import torch, sys
sys.path.insert(0, '.')
from medical_image_analysis_b200 import selective_scan_cuda_oflex as oflex
from oracle import ss_ref_c
cells = [(2048, torch.float32, True, True, True, 2), (2048, torch.float16, True, True, True, 2), (4096, torch.float32, True, True, True, 2),
         (4096, torch.float16, True, True, True, 2), (4096, torch.bfloat16, True, False, False, 1), (4096, torch.bfloat16, True, True, True, 2)]
for seqlen, itype, hb, sp, hD, groups in cells:
    torch.random.manual_seed(0)
    batch, dim, dim1, dstate = 2, 768, 24, 1
    A = -0.5 * torch.rand(dim, dstate); B = torch.randn(batch, groups, dstate, seqlen).to(itype); C = torch.randn(batch, groups, dstate, seqlen).to(itype)
    D = torch.randn(dim) if hD else None; bias = 0.5 * torch.rand(dim1) if hb else None; u = torch.randn(batch, dim, seqlen).to(itype)
    delta = (0.5 * torch.rand(batch, dim1, seqlen)).to(itype); g = torch.randn(batch, dim, seqlen).to(itype)
    cu = lambda t: None if t is None else t.cuda()
    out, x = oflex.fwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(bias), sp, 1, True)
    res = oflex.bwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(bias), cu(g).float(), x, sp, 1)
    ref = ss_ref_c.bwd(u, delta, A, B, C, D, None, bias, g, sp)
    dA = res[2].float().cpu(); r = ref["dA"]
    bad = ~torch.isclose(dA, r, rtol=1e-3, atol=5e-3)
    for i in bad.nonzero()[:, 0].tolist():
        print(seqlen, itype, 'row', i, 'A %.4e' % A[i, 0].item(), 'got %.6e ref %.6e err %.3e' % (dA[i, 0].item(), r[i, 0].item(), (dA[i, 0] - r[i, 0]).abs().item()), 'max|dA| %.3e' % r.abs().max().item())
