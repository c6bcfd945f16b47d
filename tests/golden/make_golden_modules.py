"""Module-level golden vectors produced by EXECUTING THE REFERENCE'S OWN MODULES on CPU (build container only).

    python tests/golden/make_golden_modules.py        # needs /root/reference; writes tests/golden/modules_*.npz

What runs unchanged from /root/reference (imported, not copied):
  * R2GenCSR/VMamba/classification/models/vmamba.py: cross_selective_scan, SS2D (forwardv0 / v2 family / xv family),
    VSSBlock, VSSM, Backbone_VSSM, the CrossScan ablations;
  * CXPMRG_Bench_MambaXray_VL/arm/Finetuning/mamba_simple.py + models_mamba.py: Mamba (bimamba none / v3, fast and slow
    path), Block, ARM (the arm_base_pz16 recipe at toy width);
  * HD_Xray_Pretrain_MAE/pretrain/patch_embed.py (SmallPatchEmbed) and models/mae.py (MaskedAutoencoderViT).

What is stubbed so that those files import and run without a GPU (stubs live in THIS file, test infrastructure):
  * timm / fvcore: DropPath (identity: drop rates are 0), trunc_normal_, lecun_normal_, to_2tuple, register_model, _cfg;
    timm.models.vision_transformer.Block / PatchEmbed = the standard pre-LN ViT block of timm 0.9.2 RESTATED here
    (third-party code absent from /root/reference; requirements.txt:93 pins timm==0.9.2) -> the MAE vectors pin the
    reference's orchestration (patch encode, masking, gather / scatter, decoder), not timm's block;
  * the CUDA extensions selective_scan_cuda_oflex / _core / selective_scan_cuda: fwd / bwd implemented with the
    reference's own `selective_scan_ref` (lifted from test_selective_scan.py:168-234) + torch autograd;
  * the Triton CrossScan / CrossMerge kernels (csm_triton.py): replaced by the reference's torch CrossScan / CrossMerge
    (vmamba.py:25-67), which the reference's own CHECKS.check_csm_triton (vmamba.py:1910-2039) asserts are identical;
    CrossScanTriton1b1 by its torch statement (same four index maps applied to four inputs);
  * mamba_ssm's fused mamba_inner_fn / mamba_inner_fn_no_out_proj (un-vendored fork): evaluated by running the
    REFERENCE'S OWN slow path (Mamba.forward with use_fast_path=False, mamba_simple.py:665-709) on a throw-away reference
    Mamba whose in_proj is the identity and whose parameters are the ones passed to the fused call -- the reference
    states that this path is what the fused op computes; selective_scan_fn = the lifted selective_scan_ref;
    causal_conv1d_fn = None (the reference's own conv1d fallback, :673).

Everything runs in fp32 on CPU with fixed seeds.  Stored per case: constructor config, state_dict, inputs, outputs,
input gradient and every parameter gradient for loss = sum(out * w) with a stored random w.
"""
import ast
import importlib
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange, repeat

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------------ stubs
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not (self.training and self.drop_prob > 0), "golden runs use drop_path 0"
        return x


def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(t, mean, std, a, b)


def lecun_normal_(t):
    fan_in = t[0].numel()
    return nn.init.trunc_normal_(t, std=math.sqrt(1.0 / fan_in) / 0.87962566103423978)


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class TimmAttention(nn.Module):
    """timm 0.9.2 vision_transformer.Attention (qkv_bias, no qk_norm, no dropout) restated."""

    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class TimmMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class TimmBlock(nn.Module):
    """timm 0.9.2 vision_transformer.Block with the arguments mae.py:64-66, 82-84 passes."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm, **_):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = TimmAttention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = TimmMlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


def lift(path, names, extra=None):
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "F": F, "rearrange": rearrange, "repeat": repeat, "nn": nn, "math": math, "np": np}
    ns.update(extra or {})
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def install_stubs():
    layers = _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=trunc_normal_, lecun_normal_=lecun_normal_, to_2tuple=to_2tuple)
    vt = _mod("timm.models.vision_transformer", VisionTransformer=object, _cfg=lambda **k: {}, _load_weights=lambda *a, **k: None,
              Block=TimmBlock, PatchEmbed=object)
    reg = _mod("timm.models.registry", register_model=lambda f: f)
    models = _mod("timm.models", layers=layers, vision_transformer=vt, registry=reg)
    _mod("timm", models=models)
    _mod("fvcore.nn", FlopCountAnalysis=None, flop_count_str=None, flop_count=None, parameter_count=None)
    _mod("fvcore")

    (selective_scan_ref,) = lift(f"{REF}/R2GenCSR/VMamba/kernels/selective_scan/test_selective_scan.py", ["selective_scan_ref"])

    def _bc4(t):
        return t

    def _fwd(u, delta, A, B, C, D, z, delta_bias, softplus, out_float):
        with torch.no_grad():
            out = selective_scan_ref(u, delta, A, B, C, D, z, delta_bias, softplus)
        return out.float() if out_float else out

    def _bwd(u, delta, A, B, C, D, z, delta_bias, dout, softplus):
        leaf = lambda t: None if t is None else t.detach().clone().requires_grad_()
        args = [leaf(t) for t in (u, delta, A, B, C, D, z, delta_bias)]
        with torch.enable_grad():
            out = selective_scan_ref(args[0], args[1], args[2], args[3], args[4], args[5], args[6], args[7], softplus)
            out.backward(dout.to(out.dtype))
        return [None if t is None else t.grad for t in args]

    class _Oflex:
        @staticmethod
        def fwd(u, delta, A, B, C, D, delta_bias, softplus, nrows, oflex=True):
            return [_fwd(u, delta, A, B, C, D, None, delta_bias, softplus, oflex), u.new_zeros(1)]

        @staticmethod
        def bwd(u, delta, A, B, C, D, delta_bias, dout, x, softplus, nrows):
            du, dd, dA, dB, dC, dD, _, db = _bwd(u, delta, A, B, C, D, None, delta_bias, dout, softplus)
            return [du, dd, dA, dB, dC, dD, db]

    class _Core:
        @staticmethod
        def fwd(u, delta, A, B, C, D, delta_bias, softplus, nrows):
            return [_fwd(u, delta, A, B, C, D, None, delta_bias, softplus, False), u.new_zeros(1)]

        bwd = _Oflex.bwd

    class _Mamba:
        @staticmethod
        def fwd(u, delta, A, B, C, D, z, delta_bias, softplus):
            return [_fwd(u, delta, A, B, C, D, z, delta_bias, softplus, False), u.new_zeros(1)]

        @staticmethod
        def bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, dz, softplus, recompute):
            du, dd, dA, dB, dC, dD, dzz, db = _bwd(u, delta, A, B, C, D, z, delta_bias, dout, softplus)
            return [du, dd, dA, dB, dC, dD, db] + ([dzz] if z is not None else [])

    sys.modules["selective_scan_cuda_oflex"] = types.SimpleNamespace(fwd=_Oflex.fwd, bwd=_Oflex.bwd)
    sys.modules["selective_scan_cuda_core"] = types.SimpleNamespace(fwd=_Core.fwd, bwd=_Core.bwd)
    sys.modules["selective_scan_cuda"] = types.SimpleNamespace(fwd=_Mamba.fwd, bwd=_Mamba.bwd)
    return selective_scan_ref


def import_vmamba():
    sys.path.insert(0, f"{REF}/R2GenCSR/VMamba/classification/models")
    vm = importlib.import_module("vmamba")

    class CrossScan1b1(torch.autograd.Function):        # torch statement of csm_triton.py:7-160 (1b1 variants)
        @staticmethod
        def forward(ctx, x):
            B, K, C, H, W = x.shape
            ctx.shape = (B, C, H, W)
            y = x.new_empty((B, 4, C, H * W))
            y[:, 0] = x[:, 0].flatten(2, 3)
            y[:, 1] = x[:, 1].transpose(2, 3).flatten(2, 3)
            y[:, 2] = x[:, 2].flatten(2, 3).flip(-1)
            y[:, 3] = x[:, 3].transpose(2, 3).flatten(2, 3).flip(-1)
            return y

        @staticmethod
        def backward(ctx, y):
            B, C, H, W = ctx.shape
            y = y.view(B, 4, C, H * W)
            x = y.new_empty((B, 4, C, H, W))
            x[:, 0] = y[:, 0].view(B, C, H, W)
            x[:, 1] = y[:, 1].view(B, C, W, H).transpose(2, 3)
            x[:, 2] = y[:, 2].flip(-1).view(B, C, H, W)
            x[:, 3] = y[:, 3].flip(-1).view(B, C, W, H).transpose(2, 3)
            return x

    vm.CrossScanTriton, vm.CrossMergeTriton, vm.CrossScanTriton1b1 = vm.CrossScan, vm.CrossMerge, CrossScan1b1
    return vm


# ------------------------------------------------------------------------------------------------ helpers
def to_np(t):
    return t.detach().float().cpu().numpy()


def run_module(m, x, seed, extra_inputs=()):
    """fwd + bwd of loss = sum(out * w); returns dict of arrays."""
    g = torch.Generator().manual_seed(seed)
    x = x.clone().requires_grad_(x.is_floating_point())
    m.zero_grad()
    out = m(x, *extra_inputs)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    res = {"x": to_np(x)}
    loss = 0.0
    for i, o in enumerate(outs):
        w = torch.randn(o.shape, generator=g)
        res[f"out{i}"] = to_np(o)
        res[f"w{i}"] = to_np(w)
        loss = loss + (o.float() * w).sum()
    loss.backward()
    if x.grad is not None:
        res["dx"] = to_np(x.grad)
    for n, p in m.named_parameters():
        res[f"param.{n}"] = to_np(p)
        if p.grad is not None:
            res[f"grad.{n}"] = to_np(p.grad)
    for n, b in m.named_buffers():
        res[f"buffer.{n}"] = to_np(b)
    return res


def randomize(m, seed, scale=1.0):
    """Make every parameter generic (the reference inits biases / D to constants, A_logs to a fixed table)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("A_logs") or "A_log" in n or "A_b_log" in n or "A_c" in n:
                p.copy_(torch.log(0.2 + 1.5 * torch.rand(p.shape, generator=g)))
            elif p.dim() == 1 and ("norm" in n and n.endswith("weight")):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.3 * torch.randn(p.shape, generator=g) * scale)
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * scale / math.sqrt(max(fan_in, 1)))


def save(store, tag, res, cfg=None):
    for k, v in res.items():
        store[f"{tag}|{k}"] = v
    if cfg is not None:
        store[f"{tag}|cfg"] = np.array(repr(cfg))


# ------------------------------------------------------------------------------------------------ VMamba side
def golden_vmamba(vm):
    store = {}
    # 1. cross_selective_scan as a function (einsum / no_einsum / dt_low_rank=False), d_state 1 and 16
    k = 0
    for N, mode in ((1, "einsum"), (16, "einsum"), (4, "no_einsum")):   # dt_low_rank=False (vmamba.py:372-376) cannot run in the
        # reference either: its grouped conv1d expects K * d_inner input channels but is handed d_inner
        g = torch.Generator().manual_seed(100 + k)
        B, D, H, W, K, R = 2, 8, 5, 6, 4, 3
        x = torch.randn(B, D, H, W, generator=g, requires_grad=True)
        if mode == "no_low_rank":
            xw = (torch.randn(K, D + 2 * N, D, generator=g) / D ** 0.5).requires_grad_()   # vmamba.py:372-376: x_proj gives D dts directly
            dtw = torch.zeros(K, D, R)                                                   # only its shape (K, D, R) is read
        else:
            xw = (torch.randn(K, R + 2 * N, D, generator=g) / D ** 0.5).requires_grad_()
            dtw = (torch.randn(K, D, R, generator=g) / R ** 0.5).requires_grad_()
        dtb = (0.3 * torch.randn(K, D, generator=g)).requires_grad_()
        A_logs = torch.log(0.2 + 1.5 * torch.rand(K * D, N, generator=g)).requires_grad_()
        Ds = torch.randn(K * D, generator=g).requires_grad_()
        norm = nn.LayerNorm(D)
        with torch.no_grad():
            norm.weight.copy_(1 + 0.2 * torch.randn(D, generator=g)); norm.bias.copy_(0.2 * torch.randn(D, generator=g))
        if mode == "no_low_rank":
            xin = x
            out = vm.cross_selective_scan(xin, xw.view(-1, D), None, dtw, dtb, A_logs, Ds, out_norm=norm, SelectiveScan=vm.SelectiveScanOflex,
                                          dt_low_rank=False)
        else:
            out = vm.cross_selective_scan(x, xw, None, dtw, dtb, A_logs, Ds, out_norm=norm, SelectiveScan=vm.SelectiveScanOflex,
                                          no_einsum=(mode == "no_einsum"))
        w = torch.randn(out.shape, generator=g)
        (out.float() * w).sum().backward()
        res = dict(x=to_np(x), xw=to_np(xw), dtw=to_np(dtw), dtb=to_np(dtb), A_logs=to_np(A_logs), Ds=to_np(Ds), nw=to_np(norm.weight),
                   nb=to_np(norm.bias), out=to_np(out), w=to_np(w), dx=to_np(x.grad), dxw=to_np(xw.grad), ddtb=to_np(dtb.grad),
                   dA_logs=to_np(A_logs.grad), dDs=to_np(Ds.grad), dnw=to_np(norm.weight.grad), dnb=to_np(norm.bias.grad))
        if mode != "no_low_rank":
            res["ddtw"] = to_np(dtw.grad)
        save(store, f"css{k}", res, dict(N=N, mode=mode))
        k += 1

    # 2. SS2D: every forward family the reference defines
    ss2d_cases = [
        ("v3noz", dict(d_state=1, ssm_ratio=2.0)), ("v2", dict(d_state=4)), ("v3", dict(d_state=16)), ("v4noz", dict(d_state=1)),
        ("v01", dict(d_state=2)), ("v1", dict(d_state=2)), ("v31d", dict(d_state=2)), ("v32d", dict(d_state=2)),
        ("v3nozact", dict(d_state=2)), ("v3none", dict(d_state=2)), ("v3sigmoid", dict(d_state=2)), ("v3softmax", dict(d_state=2)),
        ("v3dwconv3", dict(d_state=2)), ("v2no32", dict(d_state=2)),
        ("v0", dict(d_state=4)), ("v0seq", dict(d_state=4)),
        ("xv1a", dict(d_state=2)), ("xv2a", dict(d_state=2)), ("xv3a", dict(d_state=2)), ("xv6", dict(d_state=2)), ("xv7", dict(d_state=2)),
        ("v3", dict(d_state=2, channel_first=True)), ("v2", dict(d_state=2, d_conv=1, initialize="v2")),
    ]
    for i, (ft, kw) in enumerate(ss2d_cases):
        torch.manual_seed(200 + i)
        cfg = dict(d_model=8, ssm_ratio=2.0, dt_rank=3, forward_type=ft)
        cfg.update(kw)
        try:
            m = vm.SS2D(**cfg)
        except AttributeError as e:    # xv2 / xv5: `del self.dt_projs_weight` before it exists (vmamba.py:888, 903)
            print(f"SKIP ss2d{i} {ft}: the reference cannot construct this configuration ({e})")
            continue
        randomize(m, 300 + i)
        g = torch.Generator().manual_seed(400 + i)
        x = torch.randn(2, 8, 5, 6, generator=g) if cfg.get("channel_first") else torch.randn(2, 5, 6, 8, generator=g)
        try:
            save(store, f"ss2d{i}", run_module(m, x, 500 + i), cfg)
        except RuntimeError as e:      # e.g. "dwconv3": the unconditional bf16 cast (:420) meets fp32 conv weights without autocast
            print(f"SKIP ss2d{i} {ft}: the reference itself cannot run this configuration in fp32 without autocast ({str(e)[:80]})")

    # 3. VSSBlock / VSSM / Backbone_VSSM (toy widths; drop_path 0)
    torch.manual_seed(7)
    blk = vm.VSSBlock(hidden_dim=8, drop_path=0.0, ssm_d_state=2, ssm_ratio=2.0, ssm_dt_rank=3, forward_type="v3noz", mlp_ratio=2.0)
    randomize(blk, 8)
    save(store, "vssblock0", run_module(blk, torch.randn(2, 5, 6, 8, generator=torch.Generator().manual_seed(9)), 10),
         dict(hidden_dim=8, drop_path=0.0, ssm_d_state=2, ssm_ratio=2.0, ssm_dt_rank=3, forward_type="v3noz", mlp_ratio=2.0))
    torch.manual_seed(11)
    blk = vm.VSSBlock(hidden_dim=8, drop_path=0.0, ssm_d_state=2, ssm_ratio=1.0, ssm_dt_rank=2, forward_type="v2", mlp_ratio=2.0, gmlp=True,
                      post_norm=True)
    randomize(blk, 12)
    save(store, "vssblock1", run_module(blk, torch.randn(1, 4, 4, 8, generator=torch.Generator().manual_seed(13)), 14),
         dict(hidden_dim=8, drop_path=0.0, ssm_d_state=2, ssm_ratio=1.0, ssm_dt_rank=2, forward_type="v2", mlp_ratio=2.0, gmlp=True, post_norm=True))
    vssm_cfg = dict(patch_size=4, in_chans=3, depths=[1, 2], dims=[8, 16], ssm_d_state=1, ssm_ratio=2.0, ssm_dt_rank=2, forward_type="v3noz",
                    mlp_ratio=2.0, drop_path_rate=0.0, patch_norm=True, norm_layer="ln", downsample_version="v3", patchembed_version="v2")
    torch.manual_seed(15)
    net = vm.VSSM(**vssm_cfg)
    randomize(net, 16)
    img = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(17))
    save(store, "vssm0", run_module(net, img, 18), vssm_cfg)
    save(store, "vssm0g", run_module(net, img, 19, extra_inputs=(True,)), vssm_cfg)       # global_features=True (classifier pooling)
    torch.manual_seed(20)
    bb = vm.Backbone_VSSM(out_indices=(0, 1), **{k: v for k, v in vssm_cfg.items() if k != "norm_layer"})
    randomize(bb, 21)
    save(store, "backbone0", run_module(bb, img, 22), vssm_cfg)
    np.savez_compressed(os.path.join(HERE, "modules_vmamba.npz"), **store)
    print("modules_vmamba:", len(store), "arrays")


# ------------------------------------------------------------------------------------------------ ARM / Mamba side
def install_mamba_stubs(selective_scan_ref):
    """mamba_ssm / causal_conv1d stubs; the fused ops are evaluated with the reference's OWN slow path (see module docstring)."""
    ARM_DIR = f"{REF}/CXPMRG_Bench_MambaXray_VL/arm/Finetuning"
    state = {}

    def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
        return selective_scan_ref(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)

    def inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None, D=None, delta_bias=None,
                          B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
        assert B is None and C is None and delta_softplus
        RefMamba = state["Mamba"]
        d_inner, L = xz.shape[1] // 2, xz.shape[2]
        m = RefMamba(d_model=2 * d_inner, d_state=A.shape[1], d_conv=conv1d_weight.shape[-1], expand=0.5, dt_rank=delta_proj_weight.shape[1],
                     conv_bias=conv1d_bias is not None, bias=False, use_fast_path=False, bimamba_type="none")
        eye = torch.eye(2 * d_inner, dtype=xz.dtype)
        params = {"in_proj.weight": eye, "conv1d.weight": conv1d_weight, "x_proj.weight": x_proj_weight, "dt_proj.weight": delta_proj_weight,
                  "dt_proj.bias": delta_bias, "A_log": torch.log(-A), "D": D,
                  "out_proj.weight": torch.cat([torch.eye(d_inner, dtype=xz.dtype), torch.zeros(d_inner, d_inner, dtype=xz.dtype)], 0)}
        if conv1d_bias is not None:
            params["conv1d.bias"] = conv1d_bias
        out = torch.func.functional_call(m, params, (xz.transpose(1, 2),))       # the reference's slow path, unchanged
        return out[..., :d_inner].transpose(1, 2)

    def inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A, B=None, C=None, D=None,
              delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
        y = inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D, delta_bias, None, None, delta_softplus)
        return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)

    iface = _mod("mamba_ssm.ops.selective_scan_interface", selective_scan_fn=selective_scan_fn, mamba_inner_fn=inner,
                 mamba_inner_fn_no_out_proj=inner_no_out_proj, bimamba_inner_fn=None)
    ln = _mod("mamba_ssm.ops.triton.layernorm", RMSNorm=nn.LayerNorm, layer_norm_fn=None, rms_norm_fn=None)
    ssu = _mod("mamba_ssm.ops.triton.selective_state_update", selective_state_update=None)
    tri = _mod("mamba_ssm.ops.triton", layernorm=ln, selective_state_update=ssu)
    ops = _mod("mamba_ssm.ops", selective_scan_interface=iface, triton=tri)
    gen = _mod("mamba_ssm.utils.generation", GenerationMixin=object)
    hf = _mod("mamba_ssm.utils.hf", load_config_hf=None, load_state_dict_hf=None)
    _mod("mamba_ssm.utils", generation=gen, hf=hf)
    _mod("mamba_ssm", ops=ops)
    _mod("causal_conv1d", causal_conv1d_fn=None, causal_conv1d_update=None)
    sys.path.insert(0, ARM_DIR)
    ms = importlib.import_module("mamba_simple")
    state["Mamba"] = ms.Mamba
    mm = importlib.import_module("models_mamba")
    return ms, mm


def golden_arm(ms, mm):
    store = {}
    # Mamba mixer: v3 (four directions, cls token in the middle, L = 4 * 4 + 1), none (fused and slow path must agree), with bias / layer scale
    cases = [
        dict(d_model=8, d_state=4, expand=1, bimamba_type="v3", if_devide_out=True),
        dict(d_model=8, d_state=16, expand=2, bimamba_type="none"),
        dict(d_model=8, d_state=4, expand=1, bimamba_type="none", use_fast_path=False, bias=True, init_layer_scale=0.5),
        dict(d_model=12, d_state=2, expand=1, bimamba_type="v3", dt_rank=2, conv_bias=False),
    ]
    for i, cfg in enumerate(cases):
        torch.manual_seed(600 + i)
        m = ms.Mamba(**cfg)
        randomize(m, 700 + i)
        L = 17 if cfg["bimamba_type"] == "v3" else 11
        x = torch.randn(2, L, cfg["d_model"], generator=torch.Generator().manual_seed(800 + i))
        save(store, f"mamba{i}", run_module(m, x, 900 + i), cfg)
    # Block + ARM (the arm_base_pz16 recipe, models_mamba.py:398-409, at toy width; 64 x 64 image -> 4 x 4 patches + cls)
    arm_cfg = dict(img_size=64, patch_size=16, embed_dim=8, depth=2, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
                   final_pool_type="mean", if_abs_pos_embed=True, if_rope=False, if_rope_residual=False, bimamba_type="v3", if_cls_token=True,
                   if_devide_out=True, use_middle_cls_token=True, drop_path_rate=0.0, ssm_cfg=dict(d_state=4))
    torch.manual_seed(31)
    net = mm.ARM(**arm_cfg)
    randomize(net, 32)
    img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(33))
    save(store, "arm0", run_module(net, img, 34), arm_cfg)
    np.savez_compressed(os.path.join(HERE, "modules_arm.npz"), **store)
    print("modules_arm:", len(store), "arrays")


# ------------------------------------------------------------------------------------------------ MAE side
def golden_mae():
    MAE_DIR = f"{REF}/HD_Xray_Pretrain_MAE/pretrain"
    sys.path.insert(0, MAE_DIR)
    sys.path.insert(0, f"{MAE_DIR}/models")
    pe = importlib.import_module("patch_embed")
    mae = importlib.import_module("mae")
    store = {}
    # 1. SmallPatchEmbed alone (the reference's class, unchanged), toy widths, 128 x 128 -> 2 x 2 tokens
    torch.manual_seed(41)
    cfg = dict(in_chans=1, embed_dim=24, hidden_dim=16)
    m = pe.SmallPatchEmbed(**cfg)
    randomize(m, 42)
    save(store, "spe0", run_module(m, torch.randn(2, 1, 128, 128, generator=torch.Generator().manual_seed(43)), 44), cfg)
    # 2. MaskedAutoencoderViT: the reference hard-wires SmallPatchEmbed(1, 1024, 1024) and 400 tokens (mae.py:57, patch_embed.py:30);
    #    the toy model keeps the 1280 x 1280 / 400-token geometry and shrinks only the widths (patch encoder (1, 16, 8))
    mcfg = dict(patch_size=64, embed_dim=16, depth=2, num_heads=2, decoder_embed_dim=16, decoder_depth=1, decoder_num_heads=2, mlp_ratio=2.0)
    real_spe = mae.SmallPatchEmbed
    mae.SmallPatchEmbed = lambda *_a: real_spe(1, 16, 8)
    torch.manual_seed(45)
    net = mae.MaskedAutoencoderViT(**mcfg)
    mae.SmallPatchEmbed = real_spe
    randomize(net, 46)
    with torch.no_grad():                                   # fixed sin-cos tables are buffers-in-spirit: restore them
        net.initialize_weights()
        randomize(net.patch_embed, 47)
    img = torch.randn(1, 1, 1280, 1280, generator=torch.Generator().manual_seed(48))
    real_rand = torch.rand
    for mt in (0, 1):
        drawn = []

        def rec(*a, **k):
            t = real_rand(*a, **k)
            drawn.append(t.clone())
            return t
        torch.rand = rec
        try:
            net.zero_grad()
            loss, mask = net(img, mt, 0.85, 0.95)
        finally:
            torch.rand = real_rand
        g = torch.Generator().manual_seed(49 + mt)
        w = torch.randn(loss.shape, generator=g)
        (loss * w).sum().backward()
        res = {"loss": to_np(loss), "mask": to_np(mask), "w": to_np(w), "img_seed": np.array([48]),
               "img_checksum": np.array([float(img.double().sum())])}
        for i, t in enumerate(drawn):
            res[f"noise{i}"] = to_np(t)
        for n, p in net.named_parameters():
            res[f"param.{n}"] = to_np(p)
            if p.grad is not None:
                res[f"grad.{n}"] = to_np(p.grad)
        save(store, f"mae{mt}", res, dict(mcfg, patch_embed_dims=(1, 16, 8)))
    np.savez_compressed(os.path.join(HERE, "modules_mae.npz"), **store)
    print("modules_mae:", len(store), "arrays")


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    torch.set_num_threads(8)
    ssr = install_stubs()
    which = sys.argv[1:] or ["vmamba", "arm", "mae"]
    if "vmamba" in which:
        golden_vmamba(import_vmamba())
    if "arm" in which:
        golden_arm(*install_mamba_stubs(ssr))
    if "mae" in which:
        golden_mae()
