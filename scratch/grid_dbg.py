import torch, sys
sys.path.insert(0, '.')
from medical_image_analysis_b200 import selective_scan_cuda_oflex as oflex
from oracle import ss_ref_c
for seqlen in (1024, 2048, 4096):
  for itype in (torch.float32,):
    torch.random.manual_seed(0)
    batch, dim, dim1, dstate, groups = 2, 768, 24, 1, 2
    A = -0.5 * torch.rand(dim, dstate); B = torch.randn(batch, groups, dstate, seqlen).to(itype); C = torch.randn(batch, groups, dstate, seqlen).to(itype)
    D = torch.randn(dim); bias = 0.5 * torch.rand(dim1); u = torch.randn(batch, dim, seqlen).to(itype)
    delta = (0.5 * torch.rand(batch, dim1, seqlen)).to(itype); g = torch.randn(batch, dim, seqlen).to(itype)
    cu = lambda t: t.cuda()
    out, x = oflex.fwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(bias), True, 1, True)
    res = oflex.bwd(cu(u), cu(delta), cu(A), cu(B), cu(C), cu(D), cu(bias), cu(g).float(), x, True, 1)
    ref = ss_ref_c.bwd(u, delta, A, B, C, D, None, bias, g, True)
    for name, got in zip(("du","ddelta","dA","dB","dC","dD","ddelta_bias"), res):
        r = ref[name]; e = (got.float().cpu() - r).abs()
        i = e.argmax()
        print(seqlen, name, 'max abs err %.3e' % e.max().item(), 'ref at that elt %.3e' % r.flatten()[i].item(), 'ref max %.3e' % r.abs().max().item())
