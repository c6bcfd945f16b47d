import torch, sys
sys.path.insert(0, '.')
from medical_image_analysis_b200 import _lib
L_ = _lib.lib()
for (B, D, L) in ((64, 1536, 196), (64, 3072, 196), (8, 1536, 6400)):
    x = torch.randn(B, D, L, device='cuda', dtype=torch.bfloat16); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    w = torch.randn(D, 4, device='cuda'); b = torch.randn(D, device='cuda'); dw = torch.empty_like(w); db = torch.empty_like(b)
    st = torch.cuda.current_stream().cuda_stream
    def f(): return L_.mia_causal_conv1d_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, D, L, 4, 1, 2, D*L, L, D*L, L, st)
    def g(): return L_.mia_causal_conv1d_bwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), dy.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(), B, D, L, 4, 1, 2, D*L, L, D*L, L, D*L, L, st)
    for fn, nacc, name in ((f, 2, 'fwd'), (g, 3, 'bwd')):
        for _ in range(3): assert fn() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / n
        print(f"conv1d {name} B={B} D={D} L={L}: {t*1e3:.1f} us  {nacc*x.numel()*2/t/1e6:.0f} GB/s")
