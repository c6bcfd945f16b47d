"""GPU-side: time the tcgen05 GEMM (mia_gemm_tn) on the shapes of the models, next to cuBLAS (torch.matmul) for context.
Prints one JSON line per shape: TFLOP/s of both, fraction of the measured bf16 peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from medical_image_analysis_b200.gemm import gemm_tn  # noqa: E402

SHAPES = {
    "patch_conv1 (B=8 x 6400 patches, 1->1024 16x16)": (8 * 6400, 1024, 256),
    "patch_conv2 (B=8 x 400, 1024->1024 4x4)": (8 * 400, 1024, 16384),
    "patch_proj (B=8 x 400, 1x1)": (8 * 400, 1024, 1024),
    "ss2d in_proj (B=64 x 196, 384->1536)": (64 * 196, 1536, 384),
    "ss2d out_proj (B=64 x 196, 768->384)": (64 * 196, 384, 768),
    "arm in_proj (B=64 x 197, 768->1536)": (64 * 197, 1536, 768),
    "arm x_proj (B=64 x 197, 768->80)": (64 * 197, 80, 768),
    "mae qkv (B=64 x 61, 1024->3072)": (64 * 61, 3072, 1024),
    "mae fc1 (B=64 x 61, 1024->4096)": (64 * 61, 4096, 1024),
    "square 8192": (8192, 8192, 8192),
}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    peak = 1671.8
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["bf16_tflops"])
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, (M, N, K) in SHAPES.items():
        if only and only not in name:
            continue
        a = (torch.randn(M, K, device="cuda") / K ** 0.5).bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        t_mine = timed(lambda: gemm_tn(a, w, bias, 1))
        t_lib = timed(lambda: torch.relu(torch.nn.functional.linear(a, w, bias.bfloat16())))
        fl = 2.0 * M * N * K
        print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "us": t_mine * 1e3, "tflops": fl / t_mine / 1e9,
                          "frac_of_measured_bf16_peak": fl / t_mine / 1e9 / peak, "cublas_us": t_lib * 1e3,
                          "cublas_tflops": fl / t_lib / 1e9}), flush=True)


if __name__ == "__main__":
    main()
