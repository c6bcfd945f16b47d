import torch, sys, os
sys.path.insert(0, '.')
from medical_image_analysis_b200 import scan_fwd, scan_bwd
def run(B, R, G, L, N=1, iters=5):
    dt = torch.bfloat16
    u = torch.randn(B, R, L, device='cuda', dtype=dt); delta = (0.5 * torch.rand(B, R, L, device='cuda')).to(dt)
    A = -0.5 * torch.rand(R, N, device='cuda'); Bm = torch.randn(B, G, N, L, device='cuda').to(dt); Cm = torch.randn(B, G, N, L, device='cuda').to(dt)
    D = torch.randn(R, device='cuda'); bias = 0.5 * torch.rand(R, device='cuda'); dout = torch.randn(B, R, L, device='cuda').to(dt)
    res = []
    for it in range(iters + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out, x, _ = scan_fwd(u, delta, A, Bm, Cm, D, None, bias, True, False)
        e[1].record()
        g = scan_bwd(u, delta, A, Bm, Cm, D, None, bias, dout, x, None, True)
        e[2].record(); torch.cuda.synchronize()
        if it >= 2: res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    f = sum(r[0] for r in res) / len(res); b = sum(r[1] for r in res) / len(res)
    print(f"B={B} R={R} L={L} N={N}: fwd {f*1e3:.0f} us  bwd {b*1e3:.0f} us  -> {B*R*L/4/(f+b)/1e3:.2f} M patch-tok/s   (env {os.environ.get('MIA_NO_ROWS_BWD','')}{os.environ.get('MIA_NO_ROWS_FWD','')})", flush=True)
for B in (4, 8, 16): run(B, 3072, 4, 6400)
for B in (4, 16, 64): run(B, 3072, 4, 196)
