"""Developer tool (GPU): where the time of one SS2D fwd + bwd goes -- kernel table from torch.profiler (CUPTI)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    from medical_image_analysis_b200.vmamba import SS2D
    dev = "cuda"
    B, H = (64, 14) if "--big" not in sys.argv else (4, 80)
    m = SS2D(d_model=384, ssm_ratio=2.0, d_state=1, forward_type="v3noz").to(dev)
    x = torch.randn(B, H, H, 384, device=dev, requires_grad=True)

    def step():
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        y.float().sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("ms per fwd+bwd:", e0.elapsed_time(e1) / 10)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))


if __name__ == "__main__":
    main()
