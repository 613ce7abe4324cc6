"""Frozen parameter pack: every static integer quantity of a calibrated, frozen I-ViT model.

After ``freeze_model`` all quantisation parameters are static, but the reference recomputes them
on every forward (weight min/max + round: quant_modules.py:68-91; ``batch_frexp`` through NumPy +
``Decimal`` with a device->host sync: quant_utils.py:164-175).  ``export_deit`` evaluates that
scale chain ONCE, with the same fp32 / fp64 operations in the same order as the reference's
forward (torch CPU ops for everything fp32, ``kernels.dyadic_host`` for batch_frexp), and stores
the results under the reference's module names:

    <linear>.weight_integer int8 [N,K]    <linear>.bias_integer int32 [N]    <linear>.out_scale f32 [N]
    <qact>.me int32 [n,2] (m,e)           <qact>.me_res (residual branch)    <qact>.scale f32 [1]
    <norm>.bias_integer int32 [C]         <softmax|gelu>.x0 int32 [1]

It accepts either this package's graphs (``deit.py``) or the reference's own model objects
(identical attribute names), and is what ``engine.Engine`` and the CPU oracle both consume.
A pack round-trips through ``save``/``load`` (.npz), which is also how a QAT checkpoint would be
deployed.
"""
from __future__ import annotations

import json

import numpy as np
import torch

from .kernels import dyadic_host

_EPS = torch.finfo(torch.float32).eps


class Pack:
    def __init__(self, meta: dict, arrays: dict):
        self.meta = dict(meta)
        self.arrays = dict(arrays)

    def __getitem__(self, k):
        return self.arrays[k]

    def __contains__(self, k):
        return k in self.arrays

    def save(self, path: str):
        np.savez_compressed(path, __meta__=np.array(json.dumps(self.meta)), **self.arrays)

    @staticmethod
    def load(path: str) -> "Pack":
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["__meta__"]))
        return Pack(meta, {k: z[k] for k in z.files if k != "__meta__"})

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in self.arrays.values()))


# --------------------------------------------------------------------------------------------
# fp32 scale-chain helpers (torch CPU, same ops as the reference)
# --------------------------------------------------------------------------------------------
def _sym_scale(bits, mn, mx):
    """quant_utils.py:51-69"""
    n = 2 ** (bits - 1) - 1
    s = torch.max(-mn, mx) / float(n)
    return s.clamp(min=_EPS)


def _act_scale(qact) -> torch.Tensor:
    """Frozen QuantAct output scale (quant_modules.py:191-192) from its stored range."""
    mn = torch.as_tensor(qact.min_val, dtype=torch.float32).detach().cpu().reshape(-1)
    mx = torch.as_tensor(qact.max_val, dtype=torch.float32).detach().cpu().reshape(-1)
    return _sym_scale(qact.activation_bit, mn, mx).reshape(-1)[:1].clone()


def _quantize(x, scale_col, bits):
    """quant_utils.py:48,90-92 with a broadcastable scale."""
    n = 2 ** (bits - 1) - 1
    q = torch.round(1. / scale_col * x)
    return torch.clamp(q, -n - 1, n)


def _linear(arrs, name, lin, s_in: torch.Tensor):
    """QuantLinear / QuantConv2d static part (quant_modules.py:68-91, 305-323).  Returns the
    per-channel output scale s_w * s_in (fp32 [N])."""
    w = lin.weight.detach().cpu().float()
    v = w.reshape(w.shape[0], -1)
    s_w = _sym_scale(lin.weight_bit, v.min(dim=1).values, v.max(dim=1).values)
    w_q = _quantize(v, s_w.view(-1, 1), lin.weight_bit)
    out_scale = s_w * s_in
    arrs[name + ".weight_integer"] = w_q.to(torch.int8).numpy()
    if lin.bias is not None:
        b_q = _quantize(lin.bias.detach().cpu().float(), out_scale, lin.bias_bit)
        if float(b_q.abs().max()) >= 2.0 ** 31:
            raise OverflowError("%s: bias_integer does not fit int32" % name)
        arrs[name + ".bias_integer"] = b_q.to(torch.int32).numpy()
    else:
        arrs[name + ".bias_integer"] = np.zeros(w.shape[0], np.int32)
    arrs[name + ".out_scale"] = out_scale.numpy().astype(np.float32)
    return out_scale


def _me(arrs, key, s_in: torch.Tensor, s_out: torch.Tensor):
    m, e = dyadic_host(s_in.numpy().reshape(-1), np.float32(s_out.reshape(-1)[0].item()))
    arrs[key] = np.stack([m, e], axis=1).astype(np.int32)


def _qact(arrs, name, qact, s_in: torch.Tensor, s_res: torch.Tensor = None) -> torch.Tensor:
    """A requantising QuantAct (quant_modules.py:197-206): dyadic table(s) + output scale."""
    s_out = _act_scale(qact)
    arrs[name + ".scale"] = s_out.numpy().astype(np.float32)
    _me(arrs, name + ".me", s_in, s_out)
    if s_res is not None:
        _me(arrs, name + ".me_res", s_res, s_out)
    return s_out


def _layernorm(arrs, name, ln, C) -> torch.Tensor:
    """IntLayerNorm static part (quant_modules.py:354-356, 374-385)."""
    dim_sqrt = torch.sqrt(torch.tensor(C, dtype=torch.float))
    sf0 = dim_sqrt / 2 ** 30
    g = ln.weight.detach().cpu().float()
    bias_int = torch.floor((ln.bias.detach().cpu().float() / g) / sf0)
    if not torch.isfinite(bias_int).all() or float(bias_int.abs().max()) >= 2.0 ** 31:
        raise OverflowError("%s: LayerNorm bias_integer does not fit int32 (gamma ~ 0?)" % name)
    arrs[name + ".bias_integer"] = bias_int.to(torch.int32).numpy()
    out_scale = sf0 * g
    arrs[name + ".out_scale"] = out_scale.numpy().astype(np.float32)
    return out_scale


def _x0(arrs, key, s: torch.Tensor):
    """floor(-1 / s) in fp32 (quant_modules.py:414, 473)."""
    x0 = int(torch.floor(-1.0 / s.reshape(-1)[0]).item())
    arrs[key] = np.array([x0], np.int32)
    return x0


# --------------------------------------------------------------------------------------------
def export_deit(model) -> Pack:
    """Walk a calibrated + frozen DeiT/ViT (``VisionTransformer`` of deit.py or of the
    reference's vit_quant.py) in forward order (vit_quant.py:254-282, 130-143, 59-88;
    layers_quant.py:144-153, 184-196) and emit the static integer parameters."""
    A = {}
    C = int(model.embed_dim)
    blocks = list(model.blocks)
    H = int(blocks[0].attn.num_heads)
    pe = model.patch_embed
    P = int(pe.patch_size[0])
    img = int(pe.img_size[0])
    n_tok = int(pe.num_patches) + 1
    hidden = int(blocks[0].mlp.fc1.out_features)
    meta = dict(arch="deit", embed_dim=C, depth=len(blocks), num_heads=H, head_dim=C // H, n_tok=n_tok,
                patch=P, img_size=img, in_chans=int(pe.proj.in_channels), num_classes=int(model.head.out_features),
                mlp_hidden=hidden, softmax_bits=int(blocks[0].attn.int_softmax.output_bit))

    s_img = _act_scale(model.qact_input)                                   # vit_quant.py:257
    A["qact_input.scale"] = s_img.numpy().astype(np.float32)
    s_conv = _linear(A, "patch_embed.proj", pe.proj, s_img)                # layers_quant.py:190
    s_pe = _qact(A, "patch_embed.qact", pe.qact, s_conv)                   # :195 (16 bit)
    # cls token rides in the carrier unquantised; qact1 recovers z = RNE(cls / s_pe)  (vit_quant.py:259-265)
    cls = model.cls_token.detach().cpu().float().reshape(-1)
    A["cls_token_integer"] = torch.round(cls / s_pe).to(torch.int32).numpy()
    s_pos = _act_scale(model.qact_pos)                                     # :264 (input mode, 16 bit)
    A["qact_pos.scale"] = s_pos.numpy().astype(np.float32)
    pos = model.pos_embed.detach().cpu().float().reshape(n_tok, C)
    A["pos_embed_integer"] = _quantize(pos, s_pos, model.qact_pos.activation_bit).to(torch.int16).numpy()
    s_x = _qact(A, "qact1", model.qact1, s_pe, s_pos)                      # :265

    for i, blk in enumerate(blocks):
        p = "blocks.%d." % i
        s_ln = _layernorm(A, p + "norm1", blk.norm1, C)                    # vit_quant.py:131
        s = _qact(A, p + "qact1", blk.qact1, s_ln)                         # :132
        at = blk.attn
        s_qkv_acc = _linear(A, p + "attn.qkv", at.qkv, s)                  # :61
        s_qkv = _qact(A, p + "attn.qact1", at.qact1, s_qkv_acc)            # :62
        s_scores = (s_qkv * s_qkv) * at.scale                              # :70-73 (QuantMatMul :226, then * self.scale)
        s_attn = _qact(A, p + "attn.qact_attn1", at.qact_attn1, s_scores)  # :74
        _x0(A, p + "attn.int_softmax.x0", s_attn)                          # :76
        s_p = torch.tensor([1 / 2 ** (at.int_softmax.output_bit - 1)], dtype=torch.float32)   # quant_modules.py:494
        s_pv = s_p * s_qkv                                                 # :79-80
        s = _qact(A, p + "attn.qact2", at.qact2, s_pv)                     # :83
        s_proj = _linear(A, p + "attn.proj", at.proj, s)                   # :84
        s_a3 = _qact(A, p + "attn.qact3", at.qact3, s_proj)                # :85 (16 bit)
        s_x2 = _qact(A, p + "qact2", blk.qact2, s_a3, s_x)                 # :135 residual
        s_ln = _layernorm(A, p + "norm2", blk.norm2, C)                    # :137
        s = _qact(A, p + "qact3", blk.qact3, s_ln)                         # :138
        mlp = blk.mlp
        s_fc1 = _linear(A, p + "mlp.fc1", mlp.fc1, s)                      # layers_quant.py:145
        s_g = _qact(A, p + "mlp.qact_gelu", mlp.qact_gelu, s_fc1)          # :146
        _x0(A, p + "mlp.act.x0", s_g * 1.702)                              # quant_modules.py:427, 414
        s_go = s_g * torch.tensor([1 / 2 ** (mlp.act.output_bit - 1)], dtype=torch.float32)   # :440-443
        s = _qact(A, p + "mlp.qact1", mlp.qact1, s_go)                     # layers_quant.py:148
        s_fc2 = _linear(A, p + "mlp.fc2", mlp.fc2, s)                      # :150
        s_m2 = _qact(A, p + "mlp.qact2", mlp.qact2, s_fc2)                 # :151 (16 bit)
        s_x = _qact(A, p + "qact4", blk.qact4, s_m2, s_x2)                 # vit_quant.py:141 residual

    s_ln = _layernorm(A, "norm", model.norm, C)                            # :271
    s = _qact(A, "qact2", model.qact2, s_ln)                               # :273 (on the cls row)
    _linear(A, "head", model.head, s)                                      # :280
    return Pack(meta, A)


def export_swin(model) -> Pack:
    """Walk a calibrated + frozen Swin (``SwinTransformer`` of swin.py or of the reference's swin_quant.py) in forward
    order (swin_quant.py:539-564, 251-301, 121-169, 328-349; layers_quant.py:144-153, 184-196) and emit the static
    integer parameters.  Per block: the gathered relative-position bias (int8 [heads, N, N], swin_quant.py:142-147) and,
    for shifted blocks, the window mask as 0/1 flags (int8 [windows, N, N], :223-247)."""
    A = {}
    pe = model.patch_embed
    P = int(pe.patch_size[0])
    img = int(pe.img_size[0])
    C0 = int(model.embed_dim)
    if model.absolute_pos_embed is not None:
        raise NotImplementedError("export_swin: absolute position embedding (ape=True) is not used by the model zoo")
    layers = list(model.layers)
    meta = dict(arch="swin", embed_dim=C0, patch=P, img_size=img, in_chans=int(pe.proj.in_channels),
                num_classes=int(model.head.out_features), grid=int(pe.grid_size[0]),
                depths=[len(l.blocks) for l in layers], num_heads=[int(l.blocks[0].attn.num_heads) for l in layers],
                window=[int(l.blocks[0].window_size) for l in layers],
                shift=[[int(b.shift_size) for b in l.blocks] for l in layers],
                mlp_hidden=[int(l.blocks[0].mlp.fc1.out_features) for l in layers],
                softmax_bits=int(layers[0].blocks[0].attn.log_int_softmax.output_bit))

    s_img = _act_scale(model.qact_input)                                   # swin_quant.py:540
    A["qact_input.scale"] = s_img.numpy().astype(np.float32)
    s_conv = _linear(A, "patch_embed.proj", pe.proj, s_img)                # layers_quant.py:190
    s = _qact(A, "patch_embed.qact_before_norm", pe.qact_before_norm, s_conv)   # :193
    s_ln = _layernorm(A, "patch_embed.norm", pe.norm, C0)                  # :194
    s_pe = _qact(A, "patch_embed.qact", pe.qact, s_ln)                     # :195 (16 bit)
    s_x = _qact(A, "qact1", model.qact1, s_pe)                             # swin_quant.py:546

    for li, layer in enumerate(layers):
        for bi, blk in enumerate(layer.blocks):
            p = "layers.%d.blocks.%d." % (li, bi)
            C = int(blk.dim)
            s_ln = _layernorm(A, p + "norm1", blk.norm1, C)                # :256
            s = _qact(A, p + "qact1", blk.qact1, s_ln)                     # :257
            at = blk.attn
            s_qkv_acc = _linear(A, p + "attn.qkv", at.qkv, s)              # :128
            s_1 = _qact(A, p + "attn.qact1", at.qact1, s_qkv_acc)          # :129
            s_scores = (s_1 * s_1) * at.scale                              # :135-138
            s_a = _qact(A, p + "attn.qact_attn1", at.qact_attn1, s_scores)  # :140
            s_t = _act_scale(at.qact_table)                                # :142-143 (input quantisation of the table)
            A[p + "attn.qact_table.scale"] = s_t.numpy().astype(np.float32)
            table = at.relative_position_bias_table.detach().cpu().float()
            tq = _quantize(table, s_t, at.qact_table.activation_bit).to(torch.int8)
            A[p + "attn.qact_table.table_integer"] = tq.numpy()
            N = int(at.window_size[0] * at.window_size[1])
            idx = at.relative_position_index.detach().cpu().reshape(-1).long()
            A[p + "attn.bias_integer"] = tq[idx].view(N, N, -1).permute(2, 0, 1).contiguous().numpy()   # :144-147
            s_2 = _qact(A, p + "attn.qact2", at.qact2, s_a, s_t)           # :149 (bias as the identity branch)
            if blk.attn_mask is not None:
                A[p + "attn_mask"] = (blk.attn_mask.detach().cpu() != 0).to(torch.int8).numpy()   # :223-247, -100 where set
            _x0(A, p + "attn.log_int_softmax.x0", s_2)                     # :156
            s_p = torch.tensor([1 / 2 ** (at.log_int_softmax.output_bit - 1)], dtype=torch.float32)
            s_pv = s_p * s_1                                               # :161-162
            s = _qact(A, p + "attn.qact3", at.qact3, s_pv)                 # :164
            s_proj = _linear(A, p + "attn.proj", at.proj, s)               # :166
            s_a4 = _qact(A, p + "attn.qact4", at.qact4, s_proj)            # :167 (16 bit)
            s_x2 = _qact(A, p + "qact2", blk.qact2, s_a4, s_x)             # :293 residual
            s_ln = _layernorm(A, p + "norm2", blk.norm2, C)                # :295
            s = _qact(A, p + "qact3", blk.qact3, s_ln)                     # :296
            mlp = blk.mlp
            s_fc1 = _linear(A, p + "mlp.fc1", mlp.fc1, s)                  # layers_quant.py:145
            s_g = _qact(A, p + "mlp.qact_gelu", mlp.qact_gelu, s_fc1)      # :146
            _x0(A, p + "mlp.act.x0", s_g * 1.702)                          # quant_modules.py:427, 414
            s_go = s_g * torch.tensor([1 / 2 ** (mlp.act.output_bit - 1)], dtype=torch.float32)
            s = _qact(A, p + "mlp.qact1", mlp.qact1, s_go)                 # layers_quant.py:148
            s_fc2 = _linear(A, p + "mlp.fc2", mlp.fc2, s)                  # :150
            s_m2 = _qact(A, p + "mlp.qact2", mlp.qact2, s_fc2)             # :151 (16 bit)
            s_x = _qact(A, p + "qact4", blk.qact4, s_m2, s_x2)             # swin_quant.py:299 residual
        if layer.downsample is not None:
            d = "layers.%d.downsample." % li
            ds = layer.downsample
            s_ln = _layernorm(A, d + "norm", ds.norm, 4 * int(ds.dim))     # :344
            s = _qact(A, d + "qact1", ds.qact1, s_ln)                      # :345
            s_red = _linear(A, d + "reduction", ds.reduction, s)           # :346 (no bias)
            s_x = _qact(A, d + "qact2", ds.qact2, s_red)                   # :347

    Cf = int(model.num_features)
    s_ln = _layernorm(A, "norm", model.norm, Cf)                           # :552
    s = _qact(A, "qact2", model.qact2, s_ln)                               # :553
    s = _qact(A, "qact3", model.qact3, s)                                  # :555 (on the token average)
    _linear(A, "head", model.head, s)                                      # :562
    return Pack(meta, A)


def check_supported(pack: Pack):
    """Domain checks of the fused kernels, raised at freeze time rather than deep inside a launch.  Every scale the
    reference can produce is inside the domain of the general kernels (x0 in [-2^24, -1]); the fused attention kernels
    need the Shiftmax input scale >= 2^-16 (|x0| <= 65535: 8-bit scores over a range of at least 0.002)."""
    for k, v in pack.arrays.items():
        if k.endswith("int_softmax.x0") and not (-65535 <= int(v[0]) <= -1):
            raise ValueError("%s = %d outside [-65535, -1]" % (k, int(v[0])))
        if k.endswith("act.x0") and not (-(1 << 24) <= int(v[0]) <= -1):
            raise ValueError("%s = %d outside [-2^24, -1]" % (k, int(v[0])))
