import torch, sys
sys.path.insert(0, '.')
from medical_image_analysis_b200.selective_scan_interface import causal_conv1d_fn
for (B, D, L) in ((64, 1536, 196), (64, 3072, 196), (8, 1536, 6400)):
    x = torch.randn(B, D, L, device='cuda', dtype=torch.bfloat16, requires_grad=True)
    w = torch.randn(D, 4, device='cuda', requires_grad=True); b = torch.randn(D, device='cuda', requires_grad=True)
    dy = torch.randn(B, D, L, device='cuda', dtype=torch.bfloat16)
    res = []
    for it in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); y = causal_conv1d_fn(x, w, b, "silu"); e[1].record(); y.backward(dy); e[2].record(); torch.cuda.synchronize()
        if it >= 3: res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    f = min(r[0] for r in res); bw = min(r[1] for r in res); n = B * D * L * 2
    print(f"conv1d B={B} D={D} L={L}: fwd {f*1e3:.0f} us ({2*n/f/1e6:.0f} GB/s)  bwd {bw*1e3:.0f} us ({3*n/bw/1e6:.0f} GB/s)")
    import torch.nn.functional as F
    res = []
    for it in range(6):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); y = F.silu(F.conv1d(x, w.to(x.dtype).unsqueeze(1), b.to(x.dtype), padding=3, groups=D)[..., :L]); e[1].record(); y.backward(dy); e[2].record(); torch.cuda.synchronize()
        if it >= 3: res.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    print(f"   torch conv1d+silu: fwd {min(r[0] for r in res)*1e3:.0f} us  bwd {min(r[1] for r in res)*1e3:.0f} us")
