"""Host-side callers of the scan path: the XFMamba network assembled around the fused sm_100a operators.

This is the "caller side" of SURVEY.md section 8 -- needed to measure end-to-end two-view pairs/sec (BASELINE.json configs
1, 3, 4, 5).  It is NOT a port of the reference's model files: the dense layers are plain torch (cuDNN / cuBLAS do the
tensor-core work), and the three scan cores are restructured around the fused kernel:

* ``ss2d_core``   -- ``SS2Dv2.forward_corev2`` (reference models/fusion_vmamba.py:1143-1188).  The reference materialises
  ``xs = cross_scan(x)`` (4x the activation) only to feed ``x_proj``.  A 1x1 projection commutes with the route
  permutation, so here ``x_proj`` runs ONCE on the un-scanned image (one GEMM over all 4 routes), only the tiny
  ``(R + 2N)``-channel result is routed (``cross_scan_fn(..., one_by_one=True)``), ``dt_proj`` stays a grouped 1x1 conv
  and CrossScan + S6 + CrossMerge of the wide tensors are the single fused kernel ``ss2d_scan``.  ``xs`` and ``ys`` never exist.
* ``shallow_fuse_core`` -- ``ShallowFuse_SS2Dv4.forward_corev2`` (:777-845): swap kernel + S6 (K=2, N=16) + split.
* ``cross_fuse_core``   -- ``Cross_SS2Dv5.forward_corev2`` (:446-578): three fused SS2D streams sharing one parameter set,
  the two view streams reading ``Cs`` of the fused stream.

Module / parameter names follow the reference so that its ``state_dict`` loads unchanged (checkpoint compatibility):
``tests/test_model_host.py`` loads the reference's weights and reproduces its logits.
All scan operators are reached through the module-level ``OPS`` namespace (the CUDA operators; tests substitute oracle-backed
callables to check this host logic on CPU).
"""
from __future__ import annotations

import math
import types
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import conv, csm, csms6s, fusion_ops, norm, proj

OPS = types.SimpleNamespace(
    ss2d_scan=fusion_ops.ss2d_scan,
    cross_scan_fn=csm.cross_scan_fn,
    selective_scan_fn=csms6s.selective_scan_fn,
    swapping_scan=fusion_ops.SwappingScan_multiview.apply,
    swapping_merge=fusion_ops.SwappingMerge_multiview.apply,
    layer_norm_2d=norm.layer_norm_2d,
    dwconv3x3_silu=conv.dwconv3x3_silu,
    dt_proj=proj.dt_proj,
    # single-launch kernels of the fusion blocks (L <= 64, N <= 16); None = compose the operators above (CPU tests do)
    cross_ss2d_x3=fusion_ops.cross_ss2d_x3,
    swap_scan_fused=fusion_ops.swap_scan_fused,
)

VARIANTS = {   # reference net_fusionmamba.py:151-159
    "tiny": dict(depths=(2, 2, 8, 2), dims=96, drop_path_rate=0.2, ssm_ratio=1.0, hidden_dim=768),
    "small": dict(depths=(2, 2, 15, 2), dims=96, drop_path_rate=0.3, ssm_ratio=2.0, hidden_dim=768),
    "base": dict(depths=(2, 2, 15, 2), dims=128, drop_path_rate=0.6, ssm_ratio=2.0, hidden_dim=1024),
}


# ---------------------------------------------------------------------------------------------------------------------
# the three scan cores
# ---------------------------------------------------------------------------------------------------------------------
def _route_small(x, x_proj_weight, K, R, N):
    """x (B, D, H, W) -> dts_r (B, K, R, L), Bs, Cs (B, K, N, L) in scan order: x_proj on the un-scanned image, then
    route only the (R + 2N)-channel result (one_by_one cross scan)."""
    B, D, H, W = x.shape
    Cp = R + 2 * N
    z = F.conv2d(x, x_proj_weight.reshape(K * Cp, D, 1, 1))                 # (B, K*Cp, H, W), all routes in one GEMM
    zs = OPS.cross_scan_fn(z.view(B, K, Cp, H, W), True, True, True, 0)     # (B, K, Cp, L)
    return torch.split(zs, [R, N, N], dim=2)


def ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, Cs_override=None, return_Cs=False):
    """y (B, D, L) fp32 = cross_merge(S6(cross_scan(x), dt_proj(x_proj(xs)), ...)), reference forward_corev2 with scan_mode
    cross2d, delta_softplus=True, ssoflex=True."""
    B, D, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    dts_r, Bs, Cs = _route_small(x, x_proj_weight, K, R, N)
    dts = OPS.dt_proj(dts_r, dt_projs_weight)                                # (B, K*D, L)
    As = -A_logs.float().exp()
    Cs_used = Cs if Cs_override is None else Cs_override
    y = OPS.ss2d_scan(x, dts.contiguous(), As, Bs.contiguous(), Cs_used.contiguous(), Ds.float(),
                      dt_projs_bias.reshape(-1).float(), True, True)
    return (y, Cs) if return_Cs else y


def shallow_fuse_core(x, x2, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds):
    """(y, y2) each (B, D, L) fp32; reference ShallowFuse_SS2Dv4.forward_corev2 before out_norm"""
    B, D, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    L = H * W
    xs = OPS.swapping_scan(x, x2)                                            # (B, 2, D, L): the operand of the x_proj GEMM
    x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs, x_proj_weight)
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = OPS.dt_proj(dts, dt_projs_weight)
    if getattr(OPS, "swap_scan_fused", None) is not None and x.is_cuda and fusion_ops.swap_scan_fused_supported(N, L):
        # swap-gather + S6 + split in one kernel: u comes from x / x2 directly, the halves go straight to y / y2
        return OPS.swap_scan_fused(x.view(B, D, L), x2.view(B, D, L), dts, -A_logs.float().exp(), Bs.contiguous(), Cs.contiguous(),
                                   Ds.float(), dt_projs_bias.reshape(-1).float())
    ys = OPS.selective_scan_fn(xs.view(B, -1, L), dts, -A_logs.float().exp(), Bs.contiguous(),
                               Cs.contiguous(), Ds.float(), dt_projs_bias.reshape(-1).float(), True, True)
    return OPS.swapping_merge(ys.view(B, K, -1, L))


def cross_fuse_core(x, x2, x_fuse, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds):
    """(y, y2, y_fuse) each (B, D, L) fp32; the view streams use the fused stream's Cs (reference :536-538, 567-569)"""
    B, D, H, W = x.shape
    K, _, R = dt_projs_weight.shape
    N = A_logs.shape[1]
    if getattr(OPS, "cross_ss2d_x3", None) is not None and x.is_cuda and fusion_ops.cross_ss2d_x3_supported(N, H, W):
        # one launch for the three streams: x_proj / dt_proj on the three images batched, then the x3 kernel
        xcat = torch.cat([x_fuse, x, x2], dim=0)
        dts_r, Bs, Cs = _route_small(xcat, x_proj_weight, K, R, N)
        dts = OPS.dt_proj(dts_r, dt_projs_weight)                            # (3B, K*D, L)
        d3, B3 = dts.view(3, B, K * D, H * W), Bs.reshape(3, B, K, N, H * W)
        Cs_fuse = Cs.reshape(3, B, K, N, H * W)[0]
        y_fuse, y, y2 = OPS.cross_ss2d_x3([x_fuse, x, x2], [d3[0], d3[1], d3[2]], [B3[0], B3[1], B3[2]], Cs_fuse,
                                          -A_logs.float().exp(), Ds.float(), dt_projs_bias.reshape(-1).float())
        return y, y2, y_fuse
    y_fuse, Cs_fuse = ss2d_core(x_fuse, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, return_Cs=True)
    y = ss2d_core(x, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, Cs_override=Cs_fuse)
    y2 = ss2d_core(x2, x_proj_weight, dt_projs_weight, dt_projs_bias, A_logs, Ds, Cs_override=Cs_fuse)
    return y, y2, y_fuse


# ---------------------------------------------------------------------------------------------------------------------
# layers (names follow the reference's state_dict)
# ---------------------------------------------------------------------------------------------------------------------
class Linear2d(nn.Linear):
    """1x1 conv with Linear-shaped weights (reference :42-49)"""

    def forward(self, x):
        return F.conv2d(x, self.weight[:, :, None, None], self.bias)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over C of a channel-first tensor (reference :52-57), without the permute/copy round trip"""

    def forward(self, x):
        return OPS.layer_norm_2d(x, self.weight, self.bias, self.eps)


class DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1.0 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _ssm_params(mod, d_state, dt_rank, d_inner, k_group, dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4):
    """S4D-real A, unit D, dt bias = softplus^-1(U[log dt_min, log dt_max]) (reference mamba_init :289-358)"""
    std = dt_rank ** -0.5
    mod.dt_projs_weight = nn.Parameter(torch.empty(k_group, d_inner, dt_rank).uniform_(-std, std))
    dt = torch.exp(torch.rand(k_group, d_inner) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
    mod.dt_projs_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
    mod.A_logs = nn.Parameter(torch.log(torch.arange(1, d_state + 1, dtype=torch.float32)).repeat(k_group * d_inner, 1))
    mod.Ds = nn.Parameter(torch.ones(k_group * d_inner))


class SS2D(nn.Module):
    """backbone mixer, forward_type v05_noz, channel-first (reference SS2Dv2 :923-1206)"""

    def __init__(self, d_model, d_state=1, ssm_ratio=2.0, d_conv=3, conv_bias=False):
        super().__init__()
        self.d_inner = int(ssm_ratio * d_model)
        self.dt_rank = math.ceil(d_model / 16)
        self.in_proj = Linear2d(d_model, self.d_inner, bias=False)
        self.conv2d = nn.Conv2d(self.d_inner, self.d_inner, d_conv, padding=(d_conv - 1) // 2, groups=self.d_inner, bias=conv_bias)
        self.x_proj_weight = nn.Parameter(torch.empty(4, self.dt_rank + 2 * d_state, self.d_inner).uniform_(-1, 1) * self.d_inner ** -0.5)
        self.out_norm = LayerNorm2d(self.d_inner)
        self.out_proj = Linear2d(self.d_inner, d_model, bias=False)
        _ssm_params(self, d_state, self.dt_rank, self.d_inner, 4)

    def forward(self, x):
        x = OPS.dwconv3x3_silu(self.in_proj(x), self.conv2d.weight, self.conv2d.bias)
        B, D, H, W = x.shape
        y = ss2d_core(x, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
        y = self.out_norm(y.view(B, D, H, W)).to(x.dtype)
        return self.out_proj(y)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = Linear2d(dim, hidden)
        self.fc2 = Linear2d(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class VSSBlock(nn.Module):
    def __init__(self, dim, drop_path, d_state, ssm_ratio, mlp_ratio=4.0):
        super().__init__()
        self.norm = LayerNorm2d(dim)
        self.op = SS2D(dim, d_state, ssm_ratio)
        self.drop_path = DropPath(drop_path)
        self.norm2 = LayerNorm2d(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.drop_path(self.op(self.norm(x)))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class Backbone(nn.Module):
    """VMamba backbone, channel-first LN2d, patch-embed v2, downsample v3 (reference Backbone_VSSM :1653-1724)"""

    def __init__(self, depths=(2, 2, 15, 2), dims=96, drop_path_rate=0.3, ssm_ratio=2.0, d_state=1, in_chans=3):
        super().__init__()
        dims = [dims * 2 ** i for i in range(len(depths))]
        self.dims = dims
        n_blocks = sum(depths)        # stochastic-depth rates rise linearly over the blocks (reference :1385)
        dpr = [drop_path_rate * i / max(n_blocks - 1, 1) for i in range(n_blocks)]
        self.patch_embed = nn.Sequential(
            nn.Conv2d(in_chans, dims[0] // 2, 3, 2, 1), nn.Identity(), LayerNorm2d(dims[0] // 2), nn.Identity(), nn.GELU(),
            nn.Conv2d(dims[0] // 2, dims[0], 3, 2, 1), nn.Identity(), LayerNorm2d(dims[0]))
        self.layers = nn.ModuleList()
        for i, depth in enumerate(depths):
            rates = dpr[sum(depths[:i]):sum(depths[:i + 1])]
            down = nn.Sequential(nn.Identity(), nn.Conv2d(dims[i], dims[i + 1], 3, 2, 1), nn.Identity(), LayerNorm2d(dims[i + 1])) \
                if i < len(depths) - 1 else nn.Identity()
            self.layers.append(nn.Sequential(OrderedDict(
                blocks=nn.Sequential(*[VSSBlock(dims[i], rates[j], d_state, ssm_ratio) for j in range(depth)]),
                downsample=down)))
        for i in range(len(depths)):
            self.add_module(f"outnorm{i}", LayerNorm2d(dims[i]))
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, x, last_only=True):
        x = self.patch_embed(x)
        outs = []
        for i, layer in enumerate(self.layers):
            o = layer.blocks(x)
            x = layer.downsample(o)
            if not last_only or i == len(self.layers) - 1:
                outs.append(getattr(self, f"outnorm{i}")(o).contiguous())
        return outs[-1] if last_only else outs


class ShallowFuseSS2D(nn.Module):
    """reference ShallowFuse_SS2Dv4 (:693-876): channel-last in/out, K=2 swap scan, SE-style cross gating"""

    def __init__(self, d_model, d_state=16, ssm_ratio=2.0):
        super().__init__()
        self.d_inner = int(ssm_ratio * d_model)
        self.dt_rank = math.ceil(d_model / 16)
        self.in_proj = nn.Linear(d_model, self.d_inner, bias=False)
        self.conv2d = nn.Conv2d(self.d_inner, self.d_inner, 3, padding=1, groups=self.d_inner, bias=True)
        self.x_proj_weight = nn.Parameter(torch.empty(2, self.dt_rank + 2 * d_state, self.d_inner).uniform_(-1, 1) * self.d_inner ** -0.5)
        self.out_norm = nn.LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)
        _ssm_params(self, d_state, self.dt_rank, self.d_inner, 2)
        self.fc1 = nn.Sequential(nn.Linear(self.d_inner, self.d_inner // 16, bias=False), nn.SiLU(inplace=True),
                                 nn.Linear(self.d_inner // 16, self.d_inner, bias=False), nn.Sigmoid())

    def forward(self, x, x2):                                   # (B, H, W, C)
        xp = self.in_proj(x).permute(0, 3, 1, 2).contiguous()
        x2p = self.in_proj(x2).permute(0, 3, 1, 2).contiguous()
        xc = OPS.dwconv3x3_silu(xp, self.conv2d.weight, self.conv2d.bias)
        x2c = OPS.dwconv3x3_silu(x2p, self.conv2d.weight, self.conv2d.bias)
        B, D, H, W = xc.shape
        y, y2 = shallow_fuse_core(xc, x2c, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
        to_last = lambda t: t.view(B, D, H * W).transpose(1, 2).reshape(B, H, W, D)
        y, y2 = self.out_norm(to_last(y)).to(x.dtype), self.out_norm(to_last(y2)).to(x.dtype)
        g = self.fc1(xp.mean(dim=(2, 3))).view(B, 1, 1, D)      # squeeze-excite on each view, applied to the OTHER view
        g2 = self.fc1(x2p.mean(dim=(2, 3))).view(B, 1, 1, D)
        return self.out_proj(y * g2), self.out_proj(y2 * g)


class ShallowFusionBlock(nn.Module):
    """reference ShallowFusionBlock_v4 (:879-920)"""

    def __init__(self, hidden_dim, d_state=16):
        super().__init__()
        self.norm = nn.BatchNorm2d(hidden_dim)
        self.shallowfuseSS2D = ShallowFuseSS2D(hidden_dim, d_state)

    def forward(self, x1, x2):
        a, b = self.shallowfuseSS2D(self.norm(x1).permute(0, 2, 3, 1), self.norm(x2).permute(0, 2, 3, 1))
        return x1 + a.permute(0, 3, 1, 2), x2 + b.permute(0, 3, 1, 2)


class CrossSS2D(nn.Module):
    """reference Cross_SS2Dv5 (:361-610): three streams (x, x2, their mean) through ONE parameter set, gated by silu(fuse)"""

    def __init__(self, d_model, d_state=16, ssm_ratio=2.0):
        super().__init__()
        self.d_inner = int(ssm_ratio * d_model)
        self.dt_rank = math.ceil(d_model / 16)
        self.in_proj = nn.Linear(d_model, self.d_inner * 2, bias=False)      # present in the reference, never used (:399)
        self.in_proj.weight.requires_grad_(False)   # it never receives a gradient there either; frozen so DDP does not wait for it
        self.in_proj_sec = nn.Linear(d_model, self.d_inner, bias=False)
        self.conv2d = nn.Conv2d(self.d_inner, self.d_inner, 3, padding=1, groups=self.d_inner, bias=True)
        self.x_proj_weight = nn.Parameter(torch.empty(4, self.dt_rank + 2 * d_state, self.d_inner).uniform_(-1, 1) * self.d_inner ** -0.5)
        self.out_norm = nn.LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)
        _ssm_params(self, d_state, self.dt_rank, self.d_inner, 4)

    def forward(self, x, x2):                                   # (B, H, W, C)
        xf = self.in_proj_sec((x + x2) / 2)
        z = F.silu(xf)
        prep = lambda t: OPS.dwconv3x3_silu(t.permute(0, 3, 1, 2).contiguous(), self.conv2d.weight, self.conv2d.bias)
        xc, x2c, xfc = prep(self.in_proj_sec(x)), prep(self.in_proj_sec(x2)), prep(xf)
        B, D, H, W = xc.shape
        ys = cross_fuse_core(xc, x2c, xfc, self.x_proj_weight, self.dt_projs_weight, self.dt_projs_bias, self.A_logs, self.Ds)
        to_last = lambda t: t.view(B, D, H * W).transpose(1, 2).reshape(B, H, W, D)
        y, y2, yf = (self.out_norm(to_last(t)).to(x.dtype) for t in ys)
        return self.out_proj(y * z + y2 * z + yf * z)


class FusionBlock(nn.Module):
    """reference FusionBlock_v5 (:613-643)"""

    def __init__(self, hidden_dim, drop_path, d_state):
        super().__init__()
        self.norm = LayerNorm2d(hidden_dim)
        self.self_attention = CrossSS2D(hidden_dim, d_state)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x1, x2):
        x = self.self_attention(self.norm(x1).permute(0, 2, 3, 1), self.norm(x2).permute(0, 2, 3, 1))
        return x1 + x2 + self.drop_path(x).permute(0, 3, 1, 2)


class DeepFusion(nn.Module):
    """reference CSSFVSSLayer_v5 (:646-690)"""

    def __init__(self, hidden_dim, depth=1, drop_path=(0.0,), d_state=16):
        super().__init__()
        self.blocks = nn.ModuleList([FusionBlock(hidden_dim, drop_path[i], d_state) for i in range(depth)])

    def forward(self, x1, x2):
        for blk in self.blocks:
            x1 = blk(x1, x2)
        return x1


class TwoViewXFMamba(nn.Module):
    """reference TwoViewXFMambaTop (net_fusionmamba.py:141-210): shared backbone on both views -> shallow fusion -> deep
    fusion -> 1x1 conv -> pooled linear head"""

    def __init__(self, outputs=2, type="small", d_state=16, backbone=None, hidden_dim=None):
        super().__init__()
        cfg = dict(VARIANTS[type])
        hidden = cfg.pop("hidden_dim") if hidden_dim is None else hidden_dim
        cfg.pop("hidden_dim", None)
        if backbone is not None:
            cfg.update(backbone)
        self.mamba_feature_extrac = Backbone(**cfg)
        # only the last stage's feature map is consumed (net_fusionmamba.py:200-201): outnorm0..2 never receive a gradient
        # in the reference either; frozen so DDP does not wait for them
        for i in range(len(self.mamba_feature_extrac.dims) - 1):
            for p_ in getattr(self.mamba_feature_extrac, f"outnorm{i}").parameters():
                p_.requires_grad_(False)
        self.shallow_mamba_fusion = ShallowFusionBlock(hidden, d_state)
        self.fusemamba = DeepFusion(hidden, 1, (0.0,), d_state)
        self.final_conv = nn.Conv2d(hidden, hidden, 1)
        self.classifier = nn.Sequential(OrderedDict(avgpool=nn.AdaptiveAvgPool2d(1), flatten=nn.Flatten(1), head=nn.Linear(hidden, outputs)))

    def forward(self, x_a, x_b):
        B = x_a.shape[0]
        # the backbone is shared between the views: run both as ONE batch of 2B images (same weights, twice the
        # parallelism for every scan launch; the reference runs two sequential passes, net_fusionmamba.py:197-198)
        z = self.mamba_feature_extrac(torch.cat([x_a, x_b], 0).expand(-1, 3, -1, -1))
        z_a, z_b = self.shallow_mamba_fusion(z[:B], z[B:])
        z = self.final_conv(self.fusemamba(z_a, z_b))
        return self.classifier(z)
