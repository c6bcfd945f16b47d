"""bf16 paths of the module mirrors: the projections run on the tcgen05 GEMM (gemm.linear / conv2d_patch) and the results
stay within bf16 resolution of the same module evaluated in fp32 (library GEMMs) with the same weights."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_small_patch_embed_bf16_gemm_path_vs_fp32():
    from medical_image_analysis_b200.mae import SmallPatchEmbed
    torch.manual_seed(0)
    m = SmallPatchEmbed(1, 128, 64).cuda()
    x = torch.randn(2, 1, 256, 256, device="cuda")
    ref = m(x)
    xb = x.bfloat16().requires_grad_()
    out = m.bfloat16()(xb)
    assert out.dtype == torch.bfloat16 and out.shape == ref.shape == (2, 16, 128)
    assert _rel(out, ref) < 2e-2
    out.float().sum().backward()
    assert xb.grad is not None and torch.isfinite(xb.grad).all()


def test_patch_encode_gemm_equals_conv_in_bf16():
    """conv2d_patch (one GEMM over patches) vs F.conv2d on the SAME bf16 operands: differences are accumulation order only."""
    import torch.nn.functional as F
    from medical_image_analysis_b200.gemm import ACT_RELU, conv2d_patch
    torch.manual_seed(1)
    x = torch.randn(2, 8, 64, 64, device="cuda").bfloat16()
    w = (torch.randn(32, 8, 4, 4, device="cuda") / 11).bfloat16()
    b = torch.randn(32, device="cuda")
    got = conv2d_patch(x, w, b, 4, ACT_RELU)
    ref = F.relu(F.conv2d(x.float(), w.float(), b, stride=4))
    assert got.shape == ref.shape
    assert torch.allclose(got.float(), ref, rtol=1e-2, atol=1e-2)


def test_vit_block_and_mae_autocast_bf16():
    from medical_image_analysis_b200.mae import Block, MaskedAutoencoderViT
    torch.manual_seed(2)
    blk = Block(128, 4, 4.0, qkv_bias=True).cuda()
    x = torch.randn(3, 61, 128, device="cuda")
    ref = blk(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = blk(x)
    assert _rel(out, ref) < 2e-2
    net = MaskedAutoencoderViT(embed_dim=64, depth=1, num_heads=2, decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=2,
                               patch_embed_dims=(1, 64, 32)).cuda()
    img = torch.randn(1, 1, 1280, 1280, device="cuda")
    noise = torch.rand(1, 400, device="cuda")
    ref_loss, _ = net(img, 0, 0.85, 0.95, noise=noise)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, mask = net(img, 0, 0.85, 0.95, noise=noise)
        loss.mean().backward()
    assert _rel(loss, ref_loss) < 3e-2
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in net.parameters())


def test_ss2d_and_arm_mixer_autocast_bf16():
    from medical_image_analysis_b200.arm import Mamba
    from medical_image_analysis_b200.vmamba import SS2D, VSSBlock
    torch.manual_seed(3)
    m = SS2D(d_model=64, d_state=1, ssm_ratio=2.0, forward_type="v3noz").cuda()
    x = torch.randn(2, 14, 14, 64, device="cuda")
    ref = m(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(x)
        out.float().sum().backward()
    assert _rel(out, ref) < 3e-2
    blk = VSSBlock(hidden_dim=64, ssm_d_state=16, ssm_ratio=1.0, forward_type="v2", mlp_ratio=2.0).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert _rel(blk(x), blk.float()(x)) < 1.0            # runs end to end; (fp32 scan under v2 vs bf16 GEMMs)
    mx = Mamba(d_model=64, d_state=16, expand=1, bimamba_type="v3").cuda()
    h = torch.randn(2, 197, 64, device="cuda")
    ref = mx(h)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = mx(h)
        out.float().sum().backward()
    assert _rel(out, ref) < 3e-2
