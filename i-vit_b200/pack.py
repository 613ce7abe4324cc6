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
A pack round-trips through ``save``/``load`` (.npz).

``from_state_dict`` builds the same pack from a SAVED QAT checkpoint (``torch.save(model.state_dict())``,
quant_train.py:261) by the buffer names the reference's TVM converter reads (TVM_benchmark/convert_model.py:12-66
``*_integer``, :69-148 ``*scaling_factor``) -- never from ``QuantAct.min_val/max_val``, which are plain attributes that a
checkpoint does not carry (quant_modules.py:133-134).  ``export_tvm_params`` writes what that
converter's ``save_params`` / ``load_qconfig`` produce, from a pack.
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

    @staticmethod
    def from_state_dict(sd: dict, num_heads=None) -> "Pack":
        """See ``pack.from_state_dict``."""
        return from_state_dict(sd, num_heads)

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
class _SDNode:
    """A module path inside a state dict: ``node.blocks[3].attn.qkv`` -> prefix 'blocks.3.attn.qkv.'; the walkers below
    read either live modules or these nodes through the same helper functions."""

    def __init__(self, sd, prefix=""):
        object.__setattr__(self, "_sd", sd)
        object.__setattr__(self, "_prefix", prefix)

    def __getattr__(self, name):
        return _SDNode(self._sd, self._prefix + name + ".")

    def __getitem__(self, i):
        return _SDNode(self._sd, self._prefix + str(i) + ".")

    def has(self, leaf):
        return (self._prefix + leaf) in self._sd

    def get(self, leaf):
        k = self._prefix + leaf
        if k not in self._sd:
            raise KeyError("state dict has no %r (was the checkpoint saved after a forward of the quantised model?)" % k)
        return torch.as_tensor(self._sd[k]).detach().cpu()


def _sym_scale(bits, mn, mx):
    """quant_utils.py:51-69"""
    n = 2 ** (bits - 1) - 1
    s = torch.max(-mn, mx) / float(n)
    return s.clamp(min=_EPS)


def _act_scale(qact) -> torch.Tensor:
    """Frozen QuantAct output scale: from its stored range (quant_modules.py:191-192) for a live module, from the saved
    ``act_scaling_factor`` buffer (convert_model.py:72-78) for a state-dict node."""
    if isinstance(qact, _SDNode):
        s = qact.get("act_scaling_factor").float().reshape(-1)[:1].clone()
        if not (float(s[0]) > 0.0):
            raise ValueError("%sact_scaling_factor is %g: this QuantAct never ran before the checkpoint was saved" % (qact._prefix, float(s[0])))
        return s
    mn = torch.as_tensor(qact.min_val, dtype=torch.float32).detach().cpu().reshape(-1)
    mx = torch.as_tensor(qact.max_val, dtype=torch.float32).detach().cpu().reshape(-1)
    return _sym_scale(qact.activation_bit, mn, mx).reshape(-1)[:1].clone()


def _param(mod, name) -> torch.Tensor:
    """A parameter / buffer of a live module or of a state-dict node, as an fp32 CPU tensor."""
    if isinstance(mod, _SDNode):
        return mod.get(name).float()
    return getattr(mod, name).detach().cpu().float()


def _quantize(x, scale_col, bits):
    """quant_utils.py:48,90-92 with a broadcastable scale."""
    n = 2 ** (bits - 1) - 1
    q = torch.round(1. / scale_col * x)
    return torch.clamp(q, -n - 1, n)


def _linear(arrs, name, lin, s_in: torch.Tensor):
    """QuantLinear / QuantConv2d static part (quant_modules.py:68-91, 305-323).  Returns the
    per-channel output scale s_w * s_in (fp32 [N]).  State-dict node: the saved ``weight_integer`` / ``bias_integer`` /
    ``fc_scaling_factor`` | ``conv_scaling_factor`` buffers (convert_model.py:16-21, 85-88)."""
    if isinstance(lin, _SDNode):
        w_q = lin.get("weight_integer").float()
        w_q = w_q.reshape(w_q.shape[0], -1)
        if float(w_q.abs().max()) > 128 or not torch.equal(w_q, w_q.round()):
            raise ValueError("%sweight_integer is not an int8 tensor" % lin._prefix)
        s_w = lin.get("fc_scaling_factor" if lin.has("fc_scaling_factor") else "conv_scaling_factor").float().reshape(-1)
        out_scale = s_w * s_in
        arrs[name + ".weight_integer"] = w_q.to(torch.int8).numpy()
        if lin.has("bias_integer") and lin.has("bias"):
            arrs[name + ".bias_integer"] = lin.get("bias_integer").double().round().to(torch.int32).numpy()
        else:
            arrs[name + ".bias_integer"] = np.zeros(w_q.shape[0], np.int32)
        arrs[name + ".out_scale"] = out_scale.numpy().astype(np.float32)
        return out_scale
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
    """IntLayerNorm static part (quant_modules.py:354-356, 374-385); state-dict node: the saved ``bias_integer`` and
    ``norm_scaling_factor`` buffers (convert_model.py:49-60, 92)."""
    if isinstance(ln, _SDNode):
        bias_int = ln.get("bias_integer").double()
        out_scale = ln.get("norm_scaling_factor").float().reshape(-1)
        if out_scale.numel() != C:
            raise ValueError("%snorm_scaling_factor has %d entries, expected %d (saved before the first forward?)" % (ln._prefix, out_scale.numel(), C))
    else:
        dim_sqrt = torch.sqrt(torch.tensor(C, dtype=torch.float))
        sf0 = dim_sqrt / 2 ** 30
        g = ln.weight.detach().cpu().float()
        bias_int = torch.floor((ln.bias.detach().cpu().float() / g) / sf0)
        out_scale = sf0 * g
    if not torch.isfinite(bias_int).all() or float(bias_int.abs().max()) >= 2.0 ** 31:
        raise OverflowError("%s: LayerNorm bias_integer does not fit int32 (gamma ~ 0?)" % name)
    arrs[name + ".bias_integer"] = bias_int.to(torch.int32).numpy()
    arrs[name + ".out_scale"] = out_scale.numpy().astype(np.float32)
    return out_scale


def _x0(arrs, key, s: torch.Tensor):
    """floor(-1 / s) in fp32 (quant_modules.py:414, 473)."""
    x0 = int(torch.floor(-1.0 / s.reshape(-1)[0]).item())
    arrs[key] = np.array([x0], np.int32)
    return x0


# --------------------------------------------------------------------------------------------
def _attr(mod, name, default):
    """A plain attribute of a live module; ``default`` for a state-dict node (architecture constants)."""
    return default if isinstance(mod, _SDNode) else getattr(mod, name)


def export_deit(model) -> Pack:
    """Walk a calibrated + frozen DeiT/ViT (``VisionTransformer`` of deit.py or of the
    reference's vit_quant.py) in forward order (vit_quant.py:254-282, 130-143, 59-88;
    layers_quant.py:144-153, 184-196) and emit the static integer parameters."""
    C = int(model.embed_dim)
    blocks = list(model.blocks)
    H = int(blocks[0].attn.num_heads)
    pe = model.patch_embed
    meta = dict(arch="deit", embed_dim=C, depth=len(blocks), num_heads=H, head_dim=C // H, n_tok=int(pe.num_patches) + 1,
                patch=int(pe.patch_size[0]), img_size=int(pe.img_size[0]), in_chans=int(pe.proj.in_channels),
                num_classes=int(model.head.out_features), mlp_hidden=int(blocks[0].mlp.fc1.out_features),
                softmax_bits=int(blocks[0].attn.int_softmax.output_bit))
    return Pack(meta, _walk_deit(model, meta))


def _walk_deit(model, meta) -> dict:
    A = {}
    C, n_tok, H = meta["embed_dim"], meta["n_tok"], meta["num_heads"]
    pe = model.patch_embed
    s_img = _act_scale(model.qact_input)                                   # vit_quant.py:257
    A["qact_input.scale"] = s_img.numpy().astype(np.float32)
    s_conv = _linear(A, "patch_embed.proj", pe.proj, s_img)                # layers_quant.py:190
    s_pe = _qact(A, "patch_embed.qact", pe.qact, s_conv)                   # :195 (16 bit)
    # cls token rides in the carrier unquantised; qact1 recovers z = RNE(cls / s_pe)  (vit_quant.py:259-265)
    cls = _param(model, "cls_token").reshape(-1)
    A["cls_token_integer"] = torch.round(cls / s_pe).to(torch.int32).numpy()
    s_pos = _act_scale(model.qact_pos)                                     # :264 (input mode, 16 bit)
    A["qact_pos.scale"] = s_pos.numpy().astype(np.float32)
    pos = _param(model, "pos_embed").reshape(n_tok, C)
    A["pos_embed_integer"] = _quantize(pos, s_pos, _attr(model.qact_pos, "activation_bit", 16)).to(torch.int16).numpy()
    s_x = _qact(A, "qact1", model.qact1, s_pe, s_pos)                      # :265

    for i in range(meta["depth"]):
        blk = model.blocks[i]
        p = "blocks.%d." % i
        s_ln = _layernorm(A, p + "norm1", blk.norm1, C)                    # vit_quant.py:131
        s = _qact(A, p + "qact1", blk.qact1, s_ln)                         # :132
        at = blk.attn
        s_qkv_acc = _linear(A, p + "attn.qkv", at.qkv, s)                  # :61
        s_qkv = _qact(A, p + "attn.qact1", at.qact1, s_qkv_acc)            # :62
        s_scores = (s_qkv * s_qkv) * _attr(at, "scale", meta["head_dim"] ** -0.5)   # :70-73 (QuantMatMul :226, then * self.scale)
        s_attn = _qact(A, p + "attn.qact_attn1", at.qact_attn1, s_scores)  # :74
        _x0(A, p + "attn.int_softmax.x0", s_attn)                          # :76
        s_p = torch.tensor([1 / 2 ** (meta["softmax_bits"] - 1)], dtype=torch.float32)   # quant_modules.py:494
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
        s_go = s_g * torch.tensor([1 / 2 ** (_attr(mlp.act, "output_bit", 8) - 1)], dtype=torch.float32)   # :440-443
        s = _qact(A, p + "mlp.qact1", mlp.qact1, s_go)                     # layers_quant.py:148
        s_fc2 = _linear(A, p + "mlp.fc2", mlp.fc2, s)                      # :150
        s_m2 = _qact(A, p + "mlp.qact2", mlp.qact2, s_fc2)                 # :151 (16 bit)
        s_x = _qact(A, p + "qact4", blk.qact4, s_m2, s_x2)                 # vit_quant.py:141 residual

    s_ln = _layernorm(A, "norm", model.norm, C)                            # :271
    s = _qact(A, "qact2", model.qact2, s_ln)                               # :273 (on the cls row)
    _linear(A, "head", model.head, s)                                      # :280
    return A


def export_swin(model) -> Pack:
    """Walk a calibrated + frozen Swin (``SwinTransformer`` of swin.py or of the reference's swin_quant.py) in forward
    order (swin_quant.py:539-564, 251-301, 121-169, 328-349; layers_quant.py:144-153, 184-196) and emit the static
    integer parameters.  Per block: the gathered relative-position bias (int8 [heads, N, N], swin_quant.py:142-147) and,
    for shifted blocks, the window mask as 0/1 flags (int8 [windows, N, N], :223-247)."""
    pe = model.patch_embed
    if model.absolute_pos_embed is not None:
        raise NotImplementedError("export_swin: absolute position embedding (ape=True) is not used by the model zoo")
    layers = list(model.layers)
    meta = dict(arch="swin", embed_dim=int(model.embed_dim), patch=int(pe.patch_size[0]), img_size=int(pe.img_size[0]),
                in_chans=int(pe.proj.in_channels), num_classes=int(model.head.out_features), grid=int(pe.grid_size[0]),
                depths=[len(l.blocks) for l in layers], num_heads=[int(l.blocks[0].attn.num_heads) for l in layers],
                window=[int(l.blocks[0].window_size) for l in layers],
                shift=[[int(b.shift_size) for b in l.blocks] for l in layers],
                mlp_hidden=[int(l.blocks[0].mlp.fc1.out_features) for l in layers],
                softmax_bits=int(layers[0].blocks[0].attn.log_int_softmax.output_bit))
    return Pack(meta, _walk_swin(model, meta))


def _walk_swin(model, meta) -> dict:
    A = {}
    pe = model.patch_embed
    C0 = meta["embed_dim"]
    s_img = _act_scale(model.qact_input)                                   # swin_quant.py:540
    A["qact_input.scale"] = s_img.numpy().astype(np.float32)
    s_conv = _linear(A, "patch_embed.proj", pe.proj, s_img)                # layers_quant.py:190
    s = _qact(A, "patch_embed.qact_before_norm", pe.qact_before_norm, s_conv)   # :193
    s_ln = _layernorm(A, "patch_embed.norm", pe.norm, C0)                  # :194
    s_pe = _qact(A, "patch_embed.qact", pe.qact, s_ln)                     # :195 (16 bit)
    s_x = _qact(A, "qact1", model.qact1, s_pe)                             # swin_quant.py:546

    nl = len(meta["depths"])
    for li in range(nl):
        layer = model.layers[li]
        C = C0 * 2 ** li
        nH, ws = meta["num_heads"][li], meta["window"][li]
        for bi in range(meta["depths"][li]):
            blk = layer.blocks[bi]
            p = "layers.%d.blocks.%d." % (li, bi)
            s_ln = _layernorm(A, p + "norm1", blk.norm1, C)                # :256
            s = _qact(A, p + "qact1", blk.qact1, s_ln)                     # :257
            at = blk.attn
            s_qkv_acc = _linear(A, p + "attn.qkv", at.qkv, s)              # :128
            s_1 = _qact(A, p + "attn.qact1", at.qact1, s_qkv_acc)          # :129
            s_scores = (s_1 * s_1) * _attr(at, "scale", (C // nH) ** -0.5)  # :135-138
            s_a = _qact(A, p + "attn.qact_attn1", at.qact_attn1, s_scores)  # :140
            s_t = _act_scale(at.qact_table)                                # :142-143 (input quantisation of the table)
            A[p + "attn.qact_table.scale"] = s_t.numpy().astype(np.float32)
            table = _param(at, "relative_position_bias_table")
            tq = _quantize(table, s_t, _attr(at.qact_table, "activation_bit", 8)).to(torch.int8)
            A[p + "attn.qact_table.table_integer"] = tq.numpy()
            N = ws * ws
            idx = _param(at, "relative_position_index").reshape(-1).long()
            A[p + "attn.bias_integer"] = tq[idx].view(N, N, -1).permute(2, 0, 1).contiguous().numpy()   # :144-147
            s_2 = _qact(A, p + "attn.qact2", at.qact2, s_a, s_t)           # :149 (bias as the identity branch)
            if meta["shift"][li][bi] > 0:
                A[p + "attn_mask"] = (_param(blk, "attn_mask") != 0).to(torch.int8).numpy()   # :223-247, -100 where set
            _x0(A, p + "attn.log_int_softmax.x0", s_2)                     # :156
            s_p = torch.tensor([1 / 2 ** (meta["softmax_bits"] - 1)], dtype=torch.float32)
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
            s_go = s_g * torch.tensor([1 / 2 ** (_attr(mlp.act, "output_bit", 8) - 1)], dtype=torch.float32)
            s = _qact(A, p + "mlp.qact1", mlp.qact1, s_go)                 # layers_quant.py:148
            s_fc2 = _linear(A, p + "mlp.fc2", mlp.fc2, s)                  # :150
            s_m2 = _qact(A, p + "mlp.qact2", mlp.qact2, s_fc2)             # :151 (16 bit)
            s_x = _qact(A, p + "qact4", blk.qact4, s_m2, s_x2)             # swin_quant.py:299 residual
        if li + 1 < nl:
            d = "layers.%d.downsample." % li
            ds = layer.downsample
            s_ln = _layernorm(A, d + "norm", ds.norm, 4 * C)               # :344
            s = _qact(A, d + "qact1", ds.qact1, s_ln)                      # :345
            s_red = _linear(A, d + "reduction", ds.reduction, s)           # :346 (no bias)
            s_x = _qact(A, d + "qact2", ds.qact2, s_red)                   # :347

    Cf = C0 * 2 ** (nl - 1)
    s_ln = _layernorm(A, "norm", model.norm, Cf)                           # :552
    s = _qact(A, "qact2", model.qact2, s_ln)                               # :553
    s = _qact(A, "qact3", model.qact3, s)                                  # :555 (on the token average)
    _linear(A, "head", model.head, s)                                      # :562
    return A


# --------------------------------------------------------------------------------------------
# From a saved checkpoint
# --------------------------------------------------------------------------------------------
def _count(sd, fmt):
    n = 0
    while any(k.startswith(fmt % n) for k in sd):
        n += 1
    return n


def _softmax_bits(sd, key):
    v = float(torch.as_tensor(sd[key]).reshape(-1)[0])                     # 1 / 2^(bits-1), quant_modules.py:494
    bits = int(round(1 - np.log2(v))) if v > 0 else 0
    if bits not in (8, 16):
        raise ValueError("%s = %g is not 2^-(bits-1) for bits in {8, 16} (saved before the first forward?)" % (key, v))
    return bits


def from_state_dict(sd: dict, num_heads=None) -> Pack:
    """The frozen pack of a QAT checkpoint: ``sd`` is ``model.state_dict()`` of a reference (or mirror) DeiT/ViT/Swin
    saved after at least one forward of the quantised model (quant_train.py:261 saves it after validation), or the
    ``checkpoint['model']`` entry of such a file.  Everything is read by the reference's own keys
    (TVM_benchmark/convert_model.py:16-64 ``weight_integer`` / ``bias_integer``; :72-148 ``fc_scaling_factor``,
    ``conv_scaling_factor``, ``act_scaling_factor``, ``norm_scaling_factor``); the float ``weight`` tensors and the
    (unsaved) activation ranges are never consulted.

    ``num_heads`` (DeiT/ViT only; a state dict does not record it -- the reference's converter takes ``--depth`` the same
    way): int; defaults to embed_dim // 64 (every model of the reference zoo has head_dim 64).  Swin geometry is recovered
    from the relative-position tables / attention masks."""
    sd = sd.get("model", sd) if isinstance(sd, dict) and "model" in sd and "cls_token" not in sd else sd
    root = _SDNode(sd)
    conv_w = torch.as_tensor(sd["patch_embed.proj.weight_integer"])
    C0, in_chans, P = int(conv_w.shape[0]), int(conv_w.shape[1]), int(conv_w.shape[2])
    if "cls_token" in sd:
        n_tok = int(torch.as_tensor(sd["pos_embed"]).shape[1])
        grid = int(round((n_tok - 1) ** 0.5))
        depth = _count(sd, "blocks.%d.")
        H = int(num_heads) if num_heads else C0 // 64
        if C0 % H:
            raise ValueError("embed_dim %d is not divisible by num_heads %d" % (C0, H))
        meta = dict(arch="deit", embed_dim=C0, depth=depth, num_heads=H, head_dim=C0 // H, n_tok=n_tok, patch=P,
                    img_size=grid * P, in_chans=in_chans, num_classes=int(torch.as_tensor(sd["head.weight_integer"]).shape[0]),
                    mlp_hidden=int(torch.as_tensor(sd["blocks.0.mlp.fc1.weight_integer"]).shape[0]),
                    softmax_bits=_softmax_bits(sd, "blocks.0.attn.int_softmax.act_scaling_factor"))
        return Pack(meta, _walk_deit(root, meta))
    if "layers.0.blocks.0.attn.relative_position_bias_table" in sd:
        if "absolute_pos_embed" in sd:
            raise NotImplementedError("from_state_dict: absolute position embedding (ape=True) is not used by the model zoo")
        nl = _count(sd, "layers.%d.")
        depths = [_count(sd, "layers.%d.blocks." % li + "%d.") for li in range(nl)]
        heads, window, shift, hidden = [], [], [], []
        grid = None
        for li in range(nl):
            tab = torch.as_tensor(sd["layers.%d.blocks.0.attn.relative_position_bias_table" % li])
            ws = (int(round(tab.shape[0] ** 0.5)) + 1) // 2
            heads.append(int(tab.shape[1]))
            window.append(ws)
            hidden.append(int(torch.as_tensor(sd["layers.%d.blocks.0.mlp.fc1.weight_integer" % li]).shape[0]))
            sh = []
            for bi in range(depths[li]):
                mk = "layers.%d.blocks.%d.attn_mask" % (li, bi)
                masked = mk in sd and sd[mk] is not None
                sh.append(ws // 2 if masked else 0)
                if masked and li == 0 and grid is None:
                    grid = int(round(int(torch.as_tensor(sd[mk]).shape[0]) ** 0.5)) * ws
            shift.append(sh)
        if grid is None:
            raise ValueError("from_state_dict: cannot recover the token grid (no shifted block in stage 0)")
        meta = dict(arch="swin", embed_dim=C0, patch=P, img_size=grid * P, in_chans=in_chans,
                    num_classes=int(torch.as_tensor(sd["head.weight_integer"]).shape[0]), grid=grid, depths=depths,
                    num_heads=heads, window=window, shift=shift, mlp_hidden=hidden,
                    softmax_bits=_softmax_bits(sd, "layers.0.blocks.0.attn.log_int_softmax.act_scaling_factor"))
        return Pack(meta, _walk_swin(root, meta))
    raise ValueError("from_state_dict: neither a DeiT/ViT (cls_token) nor a Swin (relative_position_bias_table) state dict")


def export_tvm_params(pack: Pack, out_dir: str = None):
    """What TVM_benchmark/convert_model.py produces from a checkpoint, from a DeiT pack: ``(params, qconfig)``.

    params   the renamed dict of ``save_params`` (convert_model.py:23-64: ``embed_conv_weight`` int8 [C,3,P,P],
             ``embed_conv_bias`` int32 [1,C,1,1], ``block_%d_attn_qkv_weight`` / ``_bias``, ..., ``block_%d_norm1_bias``,
             ``norm_bias``, ``head_weight`` / ``head_bias``) -- written as ``params.npy`` (a pickled dict, like the
             reference's ``np.save``) when ``out_dir`` is given.  ``cls_token_weight`` / ``pos_embed_weight`` are the float
             tensors in the reference (it re-quantises them in Relay); the pack is integer-only, so they are emitted as
             ``integer * scale`` (``cls_token_integer * patch_embed.qact.scale``, ``pos_embed_integer * qact_pos.scale``),
             which quantise back to the same integers.
    qconfig  the scale chain of ``load_qconfig`` (:80-148) as {name: {input_scale, kernel_scale, output_scale}} with the
             reference's dictionary keys (``qconfig_embed_conv``, ``block_%d_qconfig_qkv``, ...), written as
             ``qconfig.json``."""
    if pack.meta.get("arch") != "deit":
        raise NotImplementedError("export_tvm_params: the reference's TVM converter covers DeiT/ViT only")
    m, A = pack.meta, pack.arrays
    P, C, depth = m["patch"], m["embed_dim"], m["depth"]
    params = {"embed_conv_weight": A["patch_embed.proj.weight_integer"].reshape(C, m["in_chans"], P, P).astype("int8"),
              "embed_conv_bias": A["patch_embed.proj.bias_integer"].astype("int32").reshape(1, -1, 1, 1)}
    for i in range(depth):
        for mod, tag in (("attn.qkv", "attn_qkv"), ("attn.proj", "attn_proj"), ("mlp.fc1", "mlp_fc1"), ("mlp.fc2", "mlp_fc2")):
            params["block_%d_%s_weight" % (i, tag)] = A["blocks.%d.%s.weight_integer" % (i, mod)].astype("int8")
            params["block_%d_%s_bias" % (i, tag)] = A["blocks.%d.%s.bias_integer" % (i, mod)].astype("int32")
        params["block_%d_norm1_bias" % i] = A["blocks.%d.norm1.bias_integer" % i].astype("int32")
        params["block_%d_norm2_bias" % i] = A["blocks.%d.norm2.bias_integer" % i].astype("int32")
    params["head_weight"] = A["head.weight_integer"].astype("int8")
    params["head_bias"] = A["head.bias_integer"].astype("int32")
    params["norm_bias"] = A["norm.bias_integer"].astype("int32")
    s_pe, s_pos = np.float32(A["patch_embed.qact.scale"][0]), np.float32(A["qact_pos.scale"][0])
    params["cls_token_weight"] = (A["cls_token_integer"].astype(np.float32) * s_pe).reshape(1, 1, C)
    params["pos_embed_weight"] = (A["pos_embed_integer"].astype(np.float32) * s_pos).reshape(1, m["n_tok"], C)

    def sc(key):
        return float(np.float32(A[key + ".scale"][0]))

    def lin(name, s_in):
        out = A[name + ".out_scale"].astype(np.float32)
        kern = (out / np.float32(s_in)).astype(np.float32)                 # fc_scaling_factor (out_scale = s_w * s_in)
        return {"input_scale": s_in, "kernel_scale": kern.tolist(), "output_scale": out.tolist()}

    q = {"qconfig_pos": {"output_scale": sc("qact_pos")},
         "qconfig_addpos": {"input_scale": sc("patch_embed.qact"), "input_dtype": "int16", "output_scale": sc("qact1")},
         "qconfig_embed_conv": lin("patch_embed.proj", float(np.float32(A["qact_input.scale"][0])))}
    s_p = float(np.float32(1.0 / 2 ** (m["softmax_bits"] - 1)))
    for i in range(depth):
        b = "blocks.%d." % i
        s_x = sc("qact1") if i == 0 else sc("blocks.%d.qact4" % (i - 1))
        s_qkv = sc(b + "attn.qact1")
        q["block_%d_qconfig_norm1" % i] = {"input_scale": s_x, "output_scale": A[b + "norm1.out_scale"].tolist()}
        q["block_%d_qconfig_qkv" % i] = lin(b + "attn.qkv", sc(b + "qact1"))
        q["block_%d_qconfig_matmul_1" % i] = {"input_scale": s_qkv, "output_scale": float(np.float32(s_qkv) * np.float32(s_qkv))}
        q["block_%d_qconfig_softmax" % i] = {"input_scale": sc(b + "attn.qact_attn1"), "output_scale": s_p}
        q["block_%d_qconfig_matmul_2" % i] = {"input_scale": s_p, "output_scale": float(np.float32(s_p) * np.float32(s_qkv))}
        q["block_%d_qconfig_proj" % i] = lin(b + "attn.proj", sc(b + "attn.qact2"))
        q["block_%d_qconfig_add1" % i] = {"input_scale": sc(b + "attn.qact3"), "input_dtype": "int16", "output_scale": sc(b + "qact2")}
        q["block_%d_qconfig_norm2" % i] = {"input_scale": sc(b + "qact2"), "output_scale": A[b + "norm2.out_scale"].tolist()}
        q["block_%d_qconfig_fc1" % i] = lin(b + "mlp.fc1", sc(b + "qact3"))
        s_g = sc(b + "mlp.qact_gelu")
        q["block_%d_qconfig_gelu" % i] = {"input_scale": s_g, "output_scale": float(np.float32(s_g) * np.float32(1 / 128)), "input_dtype": "int8"}
        q["block_%d_qconfig_fc2" % i] = lin(b + "mlp.fc2", sc(b + "mlp.qact1"))
        q["block_%d_qconfig_add2" % i] = {"input_scale": sc(b + "mlp.qact2"), "input_dtype": "int16", "output_scale": sc(b + "qact4")}
    q["qconfig_norm"] = {"input_scale": sc("blocks.%d.mlp.qact2" % (depth - 1)), "output_scale": A["norm.out_scale"].tolist()}   # :139-140 (its loop variable)
    q["qconfig_head"] = lin("head", sc("qact2"))
    if out_dir is not None:
        import os
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "params.npy"), params, allow_pickle=True)
        with open(os.path.join(out_dir, "qconfig.json"), "w") as f:
            json.dump(q, f)
    return params, q


def check_supported(pack: Pack):
    """Domain checks of the fused kernels, raised at freeze time rather than deep inside a launch.  Every scale the
    reference can produce is inside the domain of the general kernels (x0 in [-2^24, -1]); the fused attention kernels
    need the Shiftmax input scale >= 2^-16 (|x0| <= 65535: 8-bit scores over a range of at least 0.002)."""
    for k, v in pack.arrays.items():
        if k.endswith("int_softmax.x0") and not (-65535 <= int(v[0]) <= -1):
            raise ValueError("%s = %d outside [-65535, -1]" % (k, int(v[0])))
        if k.endswith("act.x0") and not (-(1 << 24) <= int(v[0]) <= -1):
            raise ValueError("%s = %d outside [-2^24, -1]" % (k, int(v[0])))
