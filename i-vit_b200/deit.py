"""DeiT / ViT integer-only inference graphs on this package's operator classes.

The reference's ``models/vit_quant.py`` / ``models/layers_quant.py`` run unchanged on the
operator mirror, but /root/reference does not travel to the GPU box, so the
benchmarked architectures are stated here as well.  Module / parameter / buffer NAMES are the
reference's (state-dict contract: a reference checkpoint loads with ``load_state_dict``, and
``pack.export_deit`` reads either kind of model object by these names); the call order of every
``forward`` follows the cited reference lines.  Inference only: dropout / drop-path are identity
in eval mode (vit_quant.py:78,86,134,140) and are not instantiated.
"""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

from .quantization_utils import IntGELU, IntLayerNorm, IntSoftmax, QuantAct, QuantConv2d, QuantLinear, QuantMatMul

__all__ = ["deit_tiny_patch16_224", "deit_small_patch16_224", "deit_base_patch16_224",
           "vit_base_patch16_224", "vit_large_patch16_224", "VisionTransformer"]


class Mlp(nn.Module):
    """layers_quant.py:116-153"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=IntGELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = QuantLinear(in_features, hidden_features)
        self.act = act_layer()
        self.qact1 = QuantAct()
        self.fc2 = QuantLinear(hidden_features, out_features)
        self.qact2 = QuantAct(16)
        self.qact_gelu = QuantAct()

    def forward(self, x, sf):
        x, sf = self.fc1(x, sf)                     # layers_quant.py:145
        x, sf = self.qact_gelu(x, sf)               # :146
        x, sf = self.act(x, sf)                     # :147
        x, sf = self.qact1(x, sf)                   # :148
        x, sf = self.fc2(x, sf)                     # :150
        return self.qact2(x, sf)                    # :151


class PatchEmbed(nn.Module):
    """layers_quant.py:156-196"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.norm_layer = norm_layer
        self.proj = QuantConv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        if self.norm_layer:
            self.qact_before_norm = QuantAct()
            self.norm = norm_layer(embed_dim)
        self.qact = QuantAct(16)

    def forward(self, x, sf):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x, sf = self.proj(x, sf)                    # layers_quant.py:190
        x = x.flatten(2).transpose(1, 2)            # :191
        if self.norm_layer:
            x, sf = self.qact_before_norm(x, sf)    # :193
            x, sf = self.norm(x, sf)                # :194
        return self.qact(x, sf)                     # :195


class Attention(nn.Module):
    """vit_quant.py:23-88"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = QuantLinear(dim, dim * 3, bias=qkv_bias)
        self.qact1 = QuantAct()
        self.qact_attn1 = QuantAct()
        self.qact2 = QuantAct()
        self.proj = QuantLinear(dim, dim)
        self.qact3 = QuantAct(16)
        self.qact_softmax = QuantAct()              # constructed, never called (vit_quant.py:51)
        self.int_softmax = IntSoftmax(16)
        self.matmul_1 = QuantMatMul()
        self.matmul_2 = QuantMatMul()

    def forward(self, x, sf):
        B, N, C = x.shape
        x, sf = self.qkv(x, sf)                                                         # vit_quant.py:61
        x, sf_1 = self.qact1(x, sf)                                                     # :62
        qkv = x.reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)   # :63-64
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn, sf = self.matmul_1(q, sf_1, k.transpose(-2, -1), sf_1)                    # :70-71
        attn = attn * self.scale                                                        # :72
        sf = sf * self.scale                                                            # :73
        attn, sf = self.qact_attn1(attn, sf)                                            # :74
        attn, sf = self.int_softmax(attn, sf)                                           # :76
        x, sf = self.matmul_2(attn, sf, v, sf_1)                                        # :79-80
        x = x.transpose(1, 2).reshape(B, N, C)                                          # :81
        x, sf = self.qact2(x, sf)                                                       # :83
        x, sf = self.proj(x, sf)                                                        # :84
        return self.qact3(x, sf)                                                        # :85


class Block(nn.Module):
    """vit_quant.py:91-143"""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=IntGELU, norm_layer=IntLayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.qact1 = QuantAct()
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale)
        self.qact2 = QuantAct(16)
        self.norm2 = norm_layer(dim)
        self.qact3 = QuantAct()
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)
        self.qact4 = QuantAct(16)

    def forward(self, x_1, sf_1):
        x, sf = self.norm1(x_1, sf_1)               # vit_quant.py:131
        x, sf = self.qact1(x, sf)                   # :132
        x, sf = self.attn(x, sf)                    # :133
        x_2, sf_2 = self.qact2(x, sf, x_1, sf_1)    # :135
        x, sf = self.norm2(x_2, sf_2)               # :137
        x, sf = self.qact3(x, sf)                   # :138
        x, sf = self.mlp(x, sf)                     # :139
        return self.qact4(x, sf, x_2, sf_2)         # :141


class VisionTransformer(nn.Module):
    """vit_quant.py:146-282"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=None):
        super().__init__()
        if representation_size:
            raise NotImplementedError("pre_logits representation layer (fp32 Linear + Tanh) is outside the integer path")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.depth, self.num_heads, self.mlp_ratio = depth, num_heads, mlp_ratio
        norm_layer = norm_layer or partial(IntLayerNorm, eps=1e-6)
        self.qact_input = QuantAct()
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.qact_pos = QuantAct(16)
        self.qact1 = QuantAct(16)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  act_layer=IntGELU, norm_layer=norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.qact2 = QuantAct()
        self.pre_logits = nn.Identity()
        self.head = QuantLinear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.act_out = QuantAct()                   # constructed, never called (vit_quant.py:240,281)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        B = x.shape[0]
        x, sf = self.qact_input(x)                                  # vit_quant.py:257
        x, sf = self.patch_embed(x, sf)                             # :258
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)  # :259-262 (cls token shares the scale)
        x_pos, sf_pos = self.qact_pos(self.pos_embed)               # :264
        x, sf = self.qact1(x, sf, x_pos, sf_pos)                    # :265
        for blk in self.blocks:
            x, sf = blk(x, sf)                                      # :268-269
        x, sf = self.norm(x, sf)                                    # :271
        x = x[:, 0]                                                 # :272
        x, sf = self.qact2(x, sf)                                   # :273
        return self.pre_logits(x), sf

    def forward(self, x):
        x, sf = self.forward_features(x)
        x, sf = self.head(x, sf)                                    # :280
        return x


def _vit(embed_dim, depth, num_heads, pretrained=False, **kwargs):
    if pretrained:
        raise RuntimeError("pretrained weights need network access; load a state_dict instead "
                           "(parameter names are the reference's)")
    return VisionTransformer(patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads, mlp_ratio=4,
                             qkv_bias=True, norm_layer=partial(IntLayerNorm, eps=1e-6), **kwargs)


def deit_tiny_patch16_224(pretrained=False, **kwargs):
    """vit_quant.py:285-303"""
    return _vit(192, 12, 3, pretrained, **kwargs)


def deit_small_patch16_224(pretrained=False, **kwargs):
    """vit_quant.py:306-323"""
    return _vit(384, 12, 6, pretrained, **kwargs)


def deit_base_patch16_224(pretrained=False, **kwargs):
    """vit_quant.py:326-343"""
    return _vit(768, 12, 12, pretrained, **kwargs)


def vit_base_patch16_224(pretrained=False, **kwargs):
    """vit_quant.py:346-362"""
    return _vit(768, 12, 12, pretrained, **kwargs)


def vit_large_patch16_224(pretrained=False, **kwargs):
    """vit_quant.py:365-381"""
    return _vit(1024, 24, 16, pretrained, **kwargs)
