"""Swin Transformer integer-only inference graphs on this package's operator classes.

Mirror of the reference's ``models/swin_quant.py`` (module / parameter / buffer names kept, so a
reference checkpoint or calibration table applies by name); every ``forward`` follows the cited
reference lines.  Inference only (dropout / drop-path are identity).  The windowing glue
(roll, window partition / reverse, 2x2 patch merging) is plain tensor indexing on the carriers,
exactly as in the reference; all integer arithmetic runs in the sm_100a operator kernels.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .deit import Mlp, PatchEmbed
from .quantization_utils import IntGELU, IntLayerNorm, IntSoftmax, QuantAct, QuantLinear, QuantMatMul

__all__ = ["swin_tiny_patch4_window7_224", "swin_small_patch4_window7_224", "swin_base_patch4_window7_224",
           "SwinTransformer"]


def window_partition(x, window_size: int):
    """(B, H, W, C) -> (num_windows*B, ws, ws, C)   swin_quant.py:18-32"""
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size: int, H: int, W: int):
    """(num_windows*B, ws, ws, C) -> (B, H, W, C)   swin_quant.py:35-50"""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


class WindowAttention(nn.Module):
    """swin_quant.py:53-169"""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        ws0, ws1 = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws0 - 1) * (2 * ws1 - 1), num_heads))
        # pair-wise relative position index of the tokens inside a window (swin_quant.py:77-91)
        coords = torch.stack(torch.meshgrid([torch.arange(ws0), torch.arange(ws1)], indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws0 - 1
        rel[:, :, 1] += ws1 - 1
        rel[:, :, 0] *= 2 * ws1 - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = QuantLinear(dim, dim * 3, bias=qkv_bias)
        self.qact1 = QuantAct()
        self.qact_attn1 = QuantAct()
        self.qact_table = QuantAct()
        self.qact2 = QuantAct()
        self.log_int_softmax = IntSoftmax()
        self.qact3 = QuantAct()
        self.qact4 = QuantAct(16)
        self.proj = QuantLinear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.matmul_1 = QuantMatMul()
        self.matmul_2 = QuantMatMul()

    def forward(self, x, sf, mask=None):
        B_, N, C = x.shape
        x, sf = self.qkv(x, sf)                                                         # swin_quant.py:128
        x, sf_1 = self.qact1(x, sf)                                                     # :129
        qkv = x.reshape(B_, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn, sf = self.matmul_1(q, sf_1, k.transpose(-2, -1), sf_1)                    # :135-136
        attn = attn * self.scale                                                        # :137
        sf = sf * self.scale                                                            # :138
        attn, sf = self.qact_attn1(attn, sf)                                            # :140
        table_q, sf_table = self.qact_table(self.relative_position_bias_table)          # :142-143
        bias = table_q[self.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()   # :144-147
        attn, sf = self.qact2(attn, sf, bias.unsqueeze(0), sf_table)                    # :149
        if mask is not None:                                                            # :151-155
            nW = mask.shape[0]
            attn = attn.view(B_ // nW, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.num_heads, N, N)
        attn, sf = self.log_int_softmax(attn, sf)                                       # :156 / :158
        x, sf = self.matmul_2(attn, sf, v, sf_1)                                        # :161-162
        x = x.transpose(1, 2).reshape(B_, N, C)                                         # :163
        x, sf = self.qact3(x, sf)                                                       # :164
        x, sf = self.proj(x, sf)                                                        # :166
        return self.qact4(x, sf)                                                        # :167


class SwinTransformerBlock(nn.Module):
    """swin_quant.py:172-301"""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 drop=0., attn_drop=0., drop_path=0., act_layer=IntGELU, norm_layer=IntLayerNorm):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:                  # no partitioning when the window covers the map
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.qact1 = QuantAct()
        self.attn = WindowAttention(dim, window_size=(self.window_size, self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias)
        self.qact2 = QuantAct(16)
        self.norm2 = norm_layer(dim)
        self.qact3 = QuantAct()
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)
        self.qact4 = QuantAct(16)
        attn_mask = None
        if self.shift_size > 0:                                             # SW-MSA mask (swin_quant.py:223-247)
            H, W = self.input_resolution
            img_mask = torch.zeros((1, H, W, 1))
            cuts = (slice(0, -self.window_size), slice(-self.window_size, -self.shift_size), slice(-self.shift_size, None))
            cnt = 0
            for h in cuts:
                for w in cuts:
                    img_mask[:, h, w, :] = cnt
                    cnt += 1
            mw = window_partition(img_mask, self.window_size).view(-1, self.window_size * self.window_size)
            attn_mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            attn_mask = attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))
        self.register_buffer("attn_mask", attn_mask)

    def forward(self, x_1, sf_1):
        H, W = self.input_resolution
        B, L, C = x_1.shape
        assert L == H * W, "input feature has wrong size"
        x, sf = self.norm1(x_1, sf_1)                                       # swin_quant.py:256
        x, sf = self.qact1(x, sf)                                           # :257
        x = x.view(B, H, W, C)
        if self.shift_size > 0:                                             # cyclic shift :261-265
            x = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        xw = window_partition(x, self.window_size).view(-1, self.window_size * self.window_size, C)   # :269-271
        aw, sf = self.attn(xw, sf, mask=self.attn_mask)                     # :275
        x = window_reverse(aw.view(-1, self.window_size, self.window_size, C), self.window_size, H, W)   # :278-281
        if self.shift_size > 0:                                             # :284-288
            x = torch.roll(x, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        x = x.view(B, H * W, C)
        x_2, sf_2 = self.qact2(x, sf, x_1, sf_1)                            # :293
        x, sf = self.norm2(x_2, sf_2)                                       # :295
        x, sf = self.qact3(x, sf)                                           # :296
        x, sf = self.mlp(x, sf)                                             # :297
        return self.qact4(x, sf, x_2, sf_2)                                 # :299


class PatchMerging(nn.Module):
    """swin_quant.py:304-349"""

    def __init__(self, input_resolution, dim, norm_layer=IntLayerNorm):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.norm = norm_layer(4 * dim)
        self.qact1 = QuantAct()
        self.reduction = QuantLinear(4 * dim, 2 * dim, bias=False)
        self.qact2 = QuantAct()

    def forward(self, x, sf):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)   # :337-341
        x = x.view(B, -1, 4 * C)
        x, sf = self.norm(x, sf)                                            # :344
        x, sf = self.qact1(x, sf)                                           # :345
        x, sf = self.reduction(x, sf)                                       # :346
        return self.qact2(x, sf)                                            # :347


class BasicLayer(nn.Module):
    """swin_quant.py:361-413"""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=IntLayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                                 shift_size=0 if (i % 2 == 0) else window_size // 2, mlp_ratio=mlp_ratio,
                                 qkv_bias=qkv_bias, act_layer=IntGELU, norm_layer=norm_layer) for i in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x, sf):
        for blk in self.blocks:
            x, sf = blk(x, sf)                                              # :409
        if self.downsample is not None:
            x, sf = self.downsample(x, sf)                                  # :411
        return x, sf


class SwinTransformer(nn.Module):
    """swin_quant.py:419-564"""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4., qkv_bias=True, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, norm_layer=IntLayerNorm, ape=False, patch_norm=True, use_checkpoint=False, **kwargs):
        super().__init__()
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.ape, self.patch_norm, self.mlp_ratio = ape, patch_norm, mlp_ratio
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.qact_input = QuantAct()
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        self.patch_grid = self.patch_embed.grid_size
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=.02)
            self.qact_pos = QuantAct(16)
        else:
            self.absolute_pos_embed = None
        self.qact1 = QuantAct(16)
        self.layers = nn.Sequential(*[
            BasicLayer(dim=int(embed_dim * 2 ** i), input_resolution=(self.patch_grid[0] // (2 ** i), self.patch_grid[1] // (2 ** i)),
                       depth=depths[i], num_heads=num_heads[i], window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                       norm_layer=norm_layer, downsample=PatchMerging if (i < self.num_layers - 1) else None)
            for i in range(self.num_layers)])
        self.norm = norm_layer(self.num_features)
        self.qact2 = QuantAct()
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.qact3 = QuantAct()
        self.head = QuantLinear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.act_out = QuantAct()                                           # constructed, never called (:518,563)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        x, sf = self.qact_input(x)                                          # swin_quant.py:540
        x, sf = self.patch_embed(x, sf)                                     # :541
        if self.absolute_pos_embed is not None:
            x_pos, sf_pos = self.qact_pos(self.absolute_pos_embed)          # :543
            x, sf = self.qact1(x, sf, x_pos, sf_pos)                        # :544
        else:
            x, sf = self.qact1(x, sf)                                       # :546
        for layer in self.layers:
            x, sf = layer(x, sf)                                            # :549-550
        x, sf = self.norm(x, sf)                                            # :552
        x, sf = self.qact2(x, sf)                                           # :553
        x = self.avgpool(x.transpose(1, 2).float())                         # :554 (on dequantised values, as the reference)
        x, sf = self.qact3(x, sf)                                           # :555
        return torch.flatten(x, 1), sf                                      # :557

    def forward(self, x):
        x, sf = self.forward_features(x)
        x, sf = self.head(x, sf)                                            # :562
        return x


def _swin(embed_dim, depths, num_heads, pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("pretrained weights need network access; load a state_dict instead")
    for k in ("quant", "calibrate", "cfg"):
        kwargs.pop(k, None)
    return SwinTransformer(patch_size=4, window_size=7, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
                           norm_layer=partial(IntLayerNorm, eps=1e-6), **kwargs)


def swin_tiny_patch4_window7_224(pretrained=False, **kwargs):
    """swin_quant.py:567-585"""
    return _swin(96, (2, 2, 6, 2), (3, 6, 12, 24), pretrained, **kwargs)


def swin_small_patch4_window7_224(pretrained=False, **kwargs):
    """swin_quant.py:588-606"""
    return _swin(96, (2, 2, 18, 2), (3, 6, 12, 24), pretrained, **kwargs)


def swin_base_patch4_window7_224(pretrained=False, **kwargs):
    """swin_quant.py:609-627"""
    return _swin(128, (2, 2, 18, 2), (4, 8, 16, 32), pretrained, **kwargs)
