"""ARM / Vim encoders of the MambaXray-VL, EMRRG and AM_MRG sub-projects on the B200 kernels.

Mirrors ``*/arm/Finetuning/mamba_simple.py`` (``Mamba``, :35-804) and ``models_mamba.py`` (``PatchEmbed`` :32-56, ``SwiGLU``
:59-83, ``Block`` :86-129, ``create_block`` :132-164, ``ARM`` :215-394, ``arm_{base,large,huge}_pz16`` :398-436): same
constructor arguments, same parameter names and shapes (``A_log``, ``D``, ``conv1d``, ``x_proj``, ``dt_proj`` and their
``_b`` / ``_c`` / ``_c_b`` / ``_d`` / ``_d_b`` direction copies), so the published MambaXray-VL checkpoints load unchanged.

What runs where: every scan is ``selective_scan_fn`` (CUDA, with the z gate), the depth-wise conv + SiLU is the CUDA
``causal_conv1d_fn``, in_proj / out_proj / SwiGLU / the patch embedding go through the tcgen05 GEMM for bf16 / fp16
activations; the four- (v3) and six- (v4) direction orchestration is the reference's, stated once over a table of
directions instead of six copies of the call.  Parity: tests/golden/modules_arm.npz holds the outputs of the reference's
own modules (its fused ops evaluated by its own slow path, see tests/golden/make_golden_modules.py).
"""
from __future__ import annotations

import math
from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gemm as _gemm
from .layernorm import RMSNorm, layer_norm_fn, rms_norm_fn
from .selective_scan_interface import (causal_conv1d_fn, causal_conv1d_update, mamba_inner_fn, mamba_inner_fn_no_out_proj, bimamba_inner_fn,
                                       selective_scan_fn)
from .vmamba import DropPath

# direction suffixes per bimamba_type: parameter sets created besides the base one (mamba_simple.py:130-389)
_EXTRA_SETS = {"none": (), "None": (), "v1": (), "v2": ("_b",), "v3": ("_b", "_c", "_c_b"), "v4": ("_b", "_c", "_c_b", "_d", "_d_b")}


def _s4d_A_log(d_inner, d_state, device=None):
    return torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1).contiguous()


class Mamba(nn.Module):
    """The ARM / Vim mixer (mamba_simple.py:35-715)."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0,
                 dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=True, layer_idx=None, device=None, dtype=None,
                 bimamba_type="none", if_devide_out=False, init_layer_scale=None):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx, self.bimamba_type, self.if_devide_out = use_fast_path, layer_idx, bimamba_type, if_devide_out
        self.init_layer_scale = init_layer_scale
        if init_layer_scale is not None:
            self.gamma = nn.Parameter(init_layer_scale * torch.ones(d_model), requires_grad=True)
        self.in_proj = nn.Linear(d_model, self.d_inner * 2, bias=bias, **fk)
        self.activation, self.act = "silu", nn.SiLU()

        def direction(sfx):
            """conv1d / x_proj / dt_proj / A_log / D of one scan direction under the reference's attribute names."""
            setattr(self, "conv1d" + sfx, nn.Conv1d(self.d_inner, self.d_inner, bias=conv_bias, kernel_size=d_conv, groups=self.d_inner,
                                                    padding=d_conv - 1, **fk))
            setattr(self, "x_proj" + sfx, nn.Linear(self.d_inner, self.dt_rank + d_state * 2, bias=False, **fk))
            setattr(self, "dt_proj" + sfx, nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk))
            A = nn.Parameter(_s4d_A_log(self.d_inner, d_state, device))
            A._no_weight_decay = True
            setattr(self, "A" + sfx + "_log", A)
            D = nn.Parameter(torch.ones(self.d_inner, device=device))
            D._no_weight_decay = True
            setattr(self, "D" + sfx, D)

        direction("")
        # only the base dt_proj gets the variance-preserving init and the inverse-softplus bias (mamba_simple.py:97-115)
        std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.dt_proj.bias._no_reinit = True
        if bimamba_type == "v1":                                      # Vim: a second A only (:131-139)
            self.A_b_log = nn.Parameter(_s4d_A_log(self.d_inner, d_state, device))
            self.A_b_log._no_weight_decay = True
        for sfx in _EXTRA_SETS.get(bimamba_type, ()):
            direction(sfx)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)

    # ------------------------------------------------------------------------------------------------------------
    def _dir_args(self, sfx):
        """positional arguments of mamba_inner_fn_no_out_proj for direction `sfx` (mamba_simple.py:450-462)."""
        dt_proj = getattr(self, "dt_proj" + sfx)
        return (getattr(self, "conv1d" + sfx).weight, getattr(self, "conv1d" + sfx).bias, getattr(self, "x_proj" + sfx).weight, dt_proj.weight,
                -torch.exp(getattr(self, "A" + sfx + "_log").float()), None, None, getattr(self, "D" + sfx).float()), \
            dict(delta_bias=dt_proj.bias.float(), delta_softplus=True)

    def _scan_dir(self, xz, sfx):
        args, kw = self._dir_args(sfx)
        return mamba_inner_fn_no_out_proj(xz, *args, **kw)

    @staticmethod
    def _col_major(t, pos):
        """(B, D, L) with the cls token at `pos`: re-order the other L - 1 tokens of the square grid column-major, keeping
        the cls token where it is (mamba_simple.py:476-482; its own inverse, :520-524)."""
        B, D, L = t.shape
        side = int(math.sqrt(L))
        cls = t[:, :, pos:pos + 1]
        grid = torch.cat([t[:, :, :pos], t[:, :, pos + 1:]], dim=-1).reshape(B, D, side, side).permute(0, 1, 3, 2).reshape(B, D, -1)
        return torch.cat((grid[:, :, :pos], cls, grid[:, :, pos:]), dim=-1)

    def _in_proj(self, h):
        """`in_proj.weight @ h` laid out (B, 2 d_inner, L) (mamba_simple.py:408-414)."""
        xz = _gemm.linear(h, self.in_proj.weight, None).transpose(1, 2)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(dtype=xz.dtype).view(1, -1, 1)
        return xz

    def forward(self, hidden_states, segmenttation_features=None, inference_params=None):
        """hidden_states (B, L, D) -> same shape; bimamba v4 returns (out, out_d).  mamba_simple.py:392-715."""
        if inference_params is not None:
            raise NotImplementedError("single-token decoding (Mamba.step / inference cache, mamba_simple.py:717-804) is not part of the "
                                      "training path (no train.py of the reference passes inference_params)")
        batch, seqlen, _ = hidden_states.shape
        xz = self._in_proj(hidden_states)
        xd = self._in_proj(segmenttation_features) if segmenttation_features is not None else None
        out_d = None
        bt = self.bimamba_type
        if self.use_fast_path and bt == "v1":                                             # :430-446
            out = bimamba_inner_fn(xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight, self.out_proj.weight,
                                   self.out_proj.bias, -torch.exp(self.A_log.float()), -torch.exp(self.A_b_log.float()), None, None,
                                   self.D.float(), delta_bias=self.dt_proj.bias.float(), delta_softplus=True)
        elif self.use_fast_path and bt in ("v3", "v4"):                                   # :447-531 / :533-648
            pos = seqlen // 2
            xc = self._col_major(xz, pos)
            out = self._scan_dir(xz, "") + self._scan_dir(xz.flip([-1]), "_b").flip([-1])
            out_c = self._col_major(self._scan_dir(xc, "_c") + self._scan_dir(xc.flip([-1]), "_c_b").flip([-1]), pos)
            if bt == "v3":
                out = _gemm.linear(((out + out_c) / 4.0).transpose(1, 2), self.out_proj.weight, self.out_proj.bias)
            else:
                out_dd = self._scan_dir(xd, "_d") + self._scan_dir(xd.flip([-1]), "_d_b").flip([-1])
                out = _gemm.linear(((out + out_c + out_dd) / 6.0).transpose(1, 2), self.out_proj.weight, self.out_proj.bias)
                out_d = _gemm.linear((out_dd / 2.0).transpose(1, 2), self.out_proj.weight, self.out_proj.bias)
        elif self.use_fast_path:                                                          # :649-663
            args, kw = self._dir_args("")
            out = mamba_inner_fn(xz, args[0], args[1], args[2], args[3], self.out_proj.weight, self.out_proj.bias, *args[4:], **kw)
        else:                                                                             # :664-709: the un-fused statement
            x, z = xz.chunk(2, dim=1)
            x = causal_conv1d_fn(x=x, weight=self.conv1d.weight.squeeze(1), bias=self.conv1d.bias, activation=self.activation)
            x_dbl = self.x_proj(x.transpose(1, 2).reshape(-1, self.d_inner))
            dt, B, C = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=-1)
            dt = (self.dt_proj.weight @ dt.t()).view(self.d_inner, batch, seqlen).transpose(0, 1)
            B = B.view(batch, seqlen, -1).transpose(1, 2).contiguous()
            C = C.view(batch, seqlen, -1).transpose(1, 2).contiguous()
            y = selective_scan_fn(x, dt, -torch.exp(self.A_log.float()), B, C, self.D.float(), z=z, delta_bias=self.dt_proj.bias.float(),
                                  delta_softplus=True)
            out = self.out_proj(y.transpose(1, 2))
        if self.init_layer_scale is not None:
            out = out * self.gamma
        return (out, out_d) if bt == "v4" else out

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        """mamba_simple.py:764-776 (shapes only; the decoding step itself is out of scope)."""
        dev = self.out_proj.weight.device
        conv_state = torch.zeros(batch_size, self.d_model * self.expand, self.d_conv, device=dev, dtype=dtype or self.conv1d.weight.dtype)
        ssm_state = torch.zeros(batch_size, self.d_model * self.expand, self.d_state, device=dev, dtype=dtype or self.dt_proj.weight.dtype)
        return conv_state, ssm_state


# ---------------------------------------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    """models_mamba.py:32-56: kernel == stride patch projection, flattened to (B, N, C).  For bf16 / fp16 inputs with
    stride == patch the conv IS a GEMM over non-overlapping patches and runs on the tcgen05 kernel."""

    def __init__(self, img_size=224, patch_size=16, stride=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        two = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        self.img_size, self.patch_size = two(img_size), two(patch_size)
        self.grid_size = ((self.img_size[0] - self.patch_size[0]) // stride + 1, (self.img_size[1] - self.patch_size[1]) // stride + 1)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=stride)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = _gemm.conv2d_patch(x, self.proj.weight, self.proj.bias, self.proj.stride)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class SwiGLU(nn.Module):
    """models_mamba.py:59-83."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.SiLU, drop=0.0, norm_layer=nn.LayerNorm, subln=False):
        super().__init__()
        out_features, hidden_features = out_features or in_features, hidden_features or in_features
        self.w1 = nn.Linear(in_features, hidden_features)
        self.w2 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.ffn_ln = norm_layer(hidden_features) if subln else nn.Identity()
        self.w3 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        hidden = self.act(_gemm.linear(x, self.w1.weight, self.w1.bias)) * _gemm.linear(x, self.w2.weight, self.w2.bias)
        return self.drop(_gemm.linear(self.ffn_ln(hidden), self.w3.weight, self.w3.bias))


class Block(nn.Module):
    """models_mamba.py:86-129: h + drop_path(mixer(LN(h))); h + drop_path(SwiGLU(LN(h))) (norm_cls / fused_add_norm are accepted
    and unused there too)."""

    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False, drop_path=0.0):
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer = mixer_cls(dim)
        self.mlp = SwiGLU(dim, dim * 4 * 2 // 3, subln=False)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)

    def forward(self, hidden_states, residual: Optional[torch.Tensor] = None, segmentation: Optional[torch.Tensor] = None, inference_params=None):
        if segmentation is None:
            hidden_states = hidden_states + self.drop_path(self.mixer(self.norm1(hidden_states), inference_params=inference_params))
            return hidden_states + self.drop_path(self.mlp(self.norm2(hidden_states)))
        feats = self.mixer(self.norm1(hidden_states), segmenttation_features=self.norm1(segmentation), inference_params=inference_params)
        hidden_states, segmentation = feats[0] + hidden_states, feats[1] + segmentation
        hidden_states = hidden_states + self.drop_path(self.mlp(self.norm2(hidden_states)))
        segmentation = segmentation + self.drop_path(self.mlp(self.norm2(segmentation)))
        return hidden_states, segmentation

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)


def create_block(d_model, ssm_cfg=None, norm_epsilon=1e-5, drop_path=0.0, rms_norm=False, residual_in_fp32=False, fused_add_norm=False,
                 layer_idx=None, device=None, dtype=None, if_bimamba=False, bimamba_type="none", if_devide_out=False, init_layer_scale=None):
    """models_mamba.py:132-164 (note expand=1)."""
    if if_bimamba:
        bimamba_type = "v1"
    fk = {"device": device, "dtype": dtype}
    mixer_cls = partial(Mamba, expand=1, layer_idx=layer_idx, bimamba_type=bimamba_type, if_devide_out=if_devide_out,
                        init_layer_scale=init_layer_scale, **(ssm_cfg or {}), **fk)
    norm_cls = partial(nn.LayerNorm if not rms_norm else RMSNorm, eps=norm_epsilon, **fk)
    block = Block(d_model, mixer_cls, norm_cls=norm_cls, drop_path=drop_path, fused_add_norm=fused_add_norm, residual_in_fp32=residual_in_fp32)
    block.layer_idx = layer_idx
    return block


def _init_weights(module, n_layer, initializer_range=0.02, rescale_prenorm_residual=True, n_residuals_per_layer=1):
    """models_mamba.py:168-198 (GPT-2 style residual rescaling)."""
    if isinstance(module, nn.Linear):
        if module.bias is not None and not getattr(module.bias, "_no_reinit", False):
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Embedding):
        nn.init.normal_(module.weight, std=initializer_range)
    if rescale_prenorm_residual:
        for name, p in module.named_parameters():
            if name in ["out_proj.weight", "fc2.weight"]:
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                with torch.no_grad():
                    p /= math.sqrt(n_residuals_per_layer * n_layer)


def segm_init_weights(m):
    """models_mamba.py:201-212."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Conv2d):
        fan_in = m.weight[0].numel()                       # timm lecun_normal_: truncated normal, variance 1 / fan_in
        nn.init.trunc_normal_(m.weight, std=math.sqrt(1.0 / fan_in) / 0.87962566103423978)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)


class ARM(nn.Module):
    """models_mamba.py:215-394: patch embed -> cls token in the MIDDLE of the sequence -> + pos_embed -> depth x Block -> LayerNorm."""

    def __init__(self, img_size=224, patch_size=16, stride=16, depth=24, embed_dim=192, channels=3, ssm_cfg=None, drop_rate=0.0,
                 drop_path_rate=0.1, norm_epsilon: float = 1e-5, rms_norm: bool = False, initializer_cfg=None, fused_add_norm=False,
                 residual_in_fp32=False, device=None, dtype=None, ft_seq_len=None, pt_hw_seq_len=14, if_bidirectional=False,
                 final_pool_type="none", if_abs_pos_embed=False, if_rope=False, if_rope_residual=False, flip_img_sequences_ratio=-1.0,
                 if_bimamba=False, bimamba_type="none", if_cls_token=False, if_devide_out=False, init_layer_scale=None,
                 use_double_cls_token=False, use_middle_cls_token=False, global_pool=False, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.residual_in_fp32, self.fused_add_norm, self.if_bidirectional = residual_in_fp32, fused_add_norm, if_bidirectional
        self.final_pool_type, self.if_abs_pos_embed, self.if_rope, self.if_rope_residual = final_pool_type, if_abs_pos_embed, if_rope, if_rope_residual
        self.flip_img_sequences_ratio, self.if_cls_token = flip_img_sequences_ratio, if_cls_token
        self.use_double_cls_token, self.use_middle_cls_token = use_double_cls_token, use_middle_cls_token
        self.num_tokens = 1 if if_cls_token else 0
        self.global_pool = global_pool
        self.d_model = self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, stride=stride, in_chans=channels, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        if if_cls_token:
            if use_double_cls_token:
                self.cls_token_head = nn.Parameter(torch.zeros(1, 1, embed_dim))
                self.cls_token_tail = nn.Parameter(torch.zeros(1, 1, embed_dim))
                self.num_tokens = 2
            else:
                self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        if if_abs_pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + self.num_tokens, embed_dim))
            self.pos_drop = nn.Dropout(p=drop_rate)
        inter_dpr = [0.0] + [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()
        self.layers = nn.ModuleList([
            create_block(embed_dim, ssm_cfg=ssm_cfg, norm_epsilon=norm_epsilon, rms_norm=rms_norm, residual_in_fp32=residual_in_fp32,
                         fused_add_norm=fused_add_norm, layer_idx=i, if_bimamba=if_bimamba, bimamba_type=bimamba_type, drop_path=inter_dpr[i],
                         if_devide_out=if_devide_out, init_layer_scale=init_layer_scale, **fk) for i in range(depth)])
        self.norm_f = nn.LayerNorm(embed_dim)
        self.patch_embed.apply(segm_init_weights)
        if if_abs_pos_embed:
            nn.init.trunc_normal_(self.pos_embed, std=0.02)
        if if_cls_token:
            for t in ((self.cls_token_head, self.cls_token_tail) if use_double_cls_token else (self.cls_token,)):
                nn.init.trunc_normal_(t, std=0.02)
        self.apply(partial(_init_weights, n_layer=depth, **(initializer_cfg or {})))

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return {i: layer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs) for i, layer in enumerate(self.layers)}

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "dist_token", "cls_token_head", "cls_token_tail"}

    def forward_features(self, x, segmentation=None, inference_params=None, if_random_cls_token_position=False, if_random_token_rank=False):
        """models_mamba.py:355-389."""
        B, M, _ = x.shape
        cls_token = self.cls_token.expand(B, -1, -1)
        pos = M // 2
        if segmentation is not None:
            mask = torch.zeros_like(x)
            for i in range(len(segmentation)):
                mask[i, segmentation[i], :] = 1
            xm = x * mask
            xm = torch.cat((xm[:, :pos, :], cls_token, xm[:, pos:, :]), dim=1)
            segmentation = self.pos_drop(xm + self.pos_embed)
        x = self.pos_drop(torch.cat((x[:, :pos, :], cls_token, x[:, pos:, :]), dim=1) + self.pos_embed)
        hidden = x
        for layer in self.layers:
            if segmentation is None:
                hidden = layer(hidden, inference_params=inference_params)
            else:
                hidden, segmentation = layer(hidden, segmentation=segmentation, inference_params=inference_params)
        return self.norm_f(hidden)

    def forward(self, x, segmentation=None, return_features=False, inference_params=None, if_random_cls_token_position=False,
                if_random_token_rank=False):
        return self.forward_features(self.patch_embed(x), segmentation, inference_params, if_random_cls_token_position=if_random_cls_token_position,
                                     if_random_token_rank=if_random_token_rank)


def _arm(embed_dim, depth, **kwargs):
    return ARM(patch_size=16, embed_dim=embed_dim, depth=depth, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
               final_pool_type="mean", if_abs_pos_embed=True, if_rope=False, if_rope_residual=False, bimamba_type="v3", if_cls_token=True,
               if_devide_out=True, use_middle_cls_token=True, **kwargs)


def arm_base_pz16(type=None, pretrained=False, **kwargs):
    """models_mamba.py:398-409 (MambaXray-VL-Base encoder; BASELINE configs[1])."""
    return _arm(768, 12, **kwargs)


def arm_large_pz16(type=None, pretrained=False, **kwargs):
    """models_mamba.py:412-423 (BASELINE configs[3])."""
    return _arm(1024, 24, **kwargs)


def arm_huge_pz16(type=None, pretrained=False, **kwargs):
    """models_mamba.py:425-436."""
    return _arm(1536, 24, **kwargs)
