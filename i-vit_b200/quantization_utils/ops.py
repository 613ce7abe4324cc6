"""Drop-in mirror of the reference's operator surface ``models/quantization_utils/quant_modules.py``
(the seven classes exported by ``models/quantization_utils/__init__.py:1``).

Same class names, constructor signatures, ``forward`` signatures and return conventions
(``(x_carrier, scaling_factor)`` tuples of fp32 "integer x scale" tensors), same buffer names
(state-dict contract used by TVM_benchmark/convert_model.py), same ``fix()/unfix()``, same
exceptions -- but every integer operation runs in the sm_100a kernels of libivit_b200.so:
carrier -> integer, integer kernel, integer -> carrier.  There is no torch / CPU fallback.

This operator-level path exists so that model code written against the reference API
(models/vit_quant.py, models/swin_quant.py, this package's ``deit``/``swin`` graphs) runs
unchanged.  It pays fp32 HBM traffic at every boundary; the fused whole-model executor
(``engine``) is the fast path and is bit-identical to it.

Operand widths (as at every call site of the reference models): QuantLinear / QuantConv2d
inputs and QuantMatMul's second operand are 8-bit; QuantMatMul's first operand may be 16-bit
(DeiT's IntSoftmax(16) output, vit_quant.py:54,79).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import kernels as K
from .primitives import fixedpoint_mul, symmetric_linear_quantization_params


def _rows(x: torch.Tensor):
    return x.reshape(-1, x.shape[-1])


class QuantLinear(nn.Linear):
    """quant_modules.py:12-97."""

    def __init__(self, in_features, out_features, bias=True, weight_bit=8, bias_bit=32, per_channel=True,
                 quant_mode='symmetric'):
        super(QuantLinear, self).__init__(in_features, out_features, bias)
        self.weight_bit = weight_bit
        self.per_channel = per_channel
        self.bias_bit = bias_bit
        self.quantize_bias = (False if bias_bit is None else True)
        self.quant_mode = quant_mode
        if self.quant_mode == "symmetric":
            pass
        elif self.quant_mode == "asymmetric":
            raise NotImplementedError("unsupported quant mode: {}".format(quant_mode))
        else:
            raise ValueError("unknown quant mode: {}".format(self.quant_mode))
        self.register_buffer('fc_scaling_factor', torch.zeros(self.out_features))
        self.register_buffer('weight_integer', torch.zeros_like(self.weight))
        if self.bias is not None:
            self.register_buffer('bias_integer', torch.zeros_like(self.bias))
        self._w_cache = None          # (weight version, int8 weights, fp32 per-row scales)

    def __repr__(self):
        s = super(QuantLinear, self).__repr__()
        return "(" + s + " weight_bit={}, quant_mode={})".format(self.weight_bit, self.quant_mode)

    def fix(self):
        pass

    def unfix(self):
        pass

    def _quantized_weight(self):
        """Per-output-channel symmetric weight quantisation (quant_modules.py:68-83).  The
        reference redoes this every forward; the weights are static, so it is cached here
        (keyed on the parameter's version counter)."""
        w = self.weight
        key = (w._version, w.data_ptr(), w.device)
        if self._w_cache is None or self._w_cache[0] != key:
            if not self.per_channel:
                raise Exception('For weight, we only support per_channel quantization.')
            with torch.no_grad():
                v = w.detach().reshape(w.shape[0], -1)
                self.min_val = v.min(axis=1).values
                self.max_val = v.max(axis=1).values
                s_w = symmetric_linear_quantization_params(self.weight_bit, self.min_val, self.max_val)
                w_q = K.quantize_f32(v, s_w, self.weight_bit, per_row=True)
            self._w_cache = (key, w_q, s_w)
            self.fc_scaling_factor = s_w
            self.weight_integer = w_q.to(torch.float32).reshape(w.shape)
        return self._w_cache[1], self._w_cache[2]

    def forward(self, x, prev_act_scaling_factor=None):
        w_q, s_w = self._quantized_weight()
        bias_scaling_factor = s_w * prev_act_scaling_factor                         # :85
        b_q = None
        if self.bias is not None:
            b_q = K.quantize_f32(self.bias.detach(), bias_scaling_factor, self.bias_bit, per_row=True,
                                 out_dtype=torch.int32)                             # :88-89
            self.bias_integer = b_q.to(torch.float32)
        else:
            self.bias_integer = None
        x_int = K.carrier_to_int(_rows(x), prev_act_scaling_factor.reshape(-1), torch.int8)   # :93-94
        out = K.gemm_i8(x_int, w_q, bias=b_q, mode="carrier", scale=bias_scaling_factor)      # :96-97
        return out.reshape(*x.shape[:-1], self.out_features), bias_scaling_factor


class QuantAct(nn.Module):
    """quant_modules.py:100-206."""

    def __init__(self, activation_bit=8, act_range_momentum=0.95, running_stat=True, per_channel=False,
                 quant_mode="symmetric"):
        super(QuantAct, self).__init__()
        self.activation_bit = activation_bit
        self.act_range_momentum = act_range_momentum
        self.running_stat = running_stat
        self.quant_mode = quant_mode
        self.per_channel = per_channel
        self.min_val = torch.zeros(1)
        self.max_val = torch.zeros(1)
        self.register_buffer('act_scaling_factor', torch.zeros(1))
        if self.quant_mode == "symmetric":
            pass
        elif self.quant_mode == "asymmetric":
            raise NotImplementedError("unsupported quant mode: {}".format(self.quant_mode))
        else:
            raise ValueError("unknown quant mode: {}".format(self.quant_mode))
        self._frozen_scale = None

    def __repr__(self):
        return "{0}(activation_bit={1}, quant_mode: {2}, Act_min: {3:.2f}, Act_max: {4:.2f})".format(
            self.__class__.__name__, self.activation_bit, self.quant_mode,
            float(torch.as_tensor(self.min_val).min()), float(torch.as_tensor(self.max_val).max()))

    def fix(self):
        """fix the activation range by setting running stat (quant_modules.py:153-157)"""
        self.running_stat = False
        self._frozen_scale = None

    def unfix(self):
        """unfix the activation range by setting running stat (quant_modules.py:159-163)"""
        self.running_stat = True
        self._frozen_scale = None

    def set_range(self, min_val: float, max_val: float):
        """Load a calibrated range (the reference keeps min_val/max_val as plain attributes that a
        checkpoint does not restore) and freeze."""
        self.min_val = torch.tensor(float(min_val), dtype=torch.float32)
        self.max_val = torch.tensor(float(max_val), dtype=torch.float32)
        self.running_stat = False
        self._frozen_scale = None

    def _scale(self, device):
        if self.running_stat or self._frozen_scale is None or self._frozen_scale.device != device:
            mn = torch.as_tensor(self.min_val, dtype=torch.float32).to(device)
            mx = torch.as_tensor(self.max_val, dtype=torch.float32).to(device)
            s = symmetric_linear_quantization_params(self.activation_bit, mn, mx).reshape(-1)[:1].contiguous()
            if self.running_stat:
                return s
            self._frozen_scale = s      # same tensor object on every frozen forward (stable data_ptr)
        return self._frozen_scale

    def forward(self, x, pre_act_scaling_factor=None, identity=None, identity_scaling_factor=None):
        with torch.no_grad():
            if self.running_stat:                                               # calibration pass :170-189
                x_act = x if identity is None else identity + x
                if len(x_act.shape) == 4:
                    x_act = x_act.permute(0, 2, 3, 1)
                v = x_act.reshape(-1, x_act.shape[-1]).transpose(0, 1)
                cur_min = v.min(axis=1).values
                cur_max = v.max(axis=1).values
                mn = torch.as_tensor(self.min_val).to(cur_min.device)
                mx = torch.as_tensor(self.max_val).to(cur_min.device)
                if torch.eq(mn, mx).all():
                    mn, mx = cur_min, cur_max
                else:
                    mn = mn * self.act_range_momentum + cur_min * (1 - self.act_range_momentum)
                    mx = mx * self.act_range_momentum + cur_max * (1 - self.act_range_momentum)
                self.max_val = mx.max()
                self.min_val = mn.min()
            s_out = self._scale(x.device)                                       # :191-192
            self.act_scaling_factor = s_out
            if pre_act_scaling_factor is None:                                  # input quantisation :194-196
                if x.dim() > 4:
                    raise NotImplementedError
                q = K.quantize_f32(x, s_out, self.activation_bit)
            else:                                                               # requant :197-202
                q = fixedpoint_mul.integer(x, pre_act_scaling_factor, self.activation_bit, self.quant_mode,
                                           s_out, identity, identity_scaling_factor)
            out = K.int_to_carrier(q, s_out)                                    # :204-206
        return out, self.act_scaling_factor


class QuantMatMul(nn.Module):
    """quant_modules.py:209-228."""

    def __init__(self):
        super(QuantMatMul, self).__init__()
        self.register_buffer('act_scaling_factor', torch.zeros(1))

    def fix(self):
        pass

    def unfix(self):
        pass

    def forward(self, A, pre_act_scaling_factor_A, B, pre_act_scaling_factor_B):
        if A.shape[:-2] != B.shape[:-2] or A.shape[-1] != B.shape[-2]:
            raise RuntimeError("QuantMatMul: incompatible shapes %s @ %s" % (tuple(A.shape), tuple(B.shape)))
        sA = pre_act_scaling_factor_A.reshape(-1)
        sB = pre_act_scaling_factor_B.reshape(-1)
        if sA.numel() != 1 or sB.numel() != 1:
            raise NotImplementedError("QuantMatMul operands carry scalar scales at every call site")
        a_int = K.carrier_to_int(A, sA, torch.int16)                            # :224
        b_int = K.carrier_to_int(B, sB, torch.int8)                             # :225
        act_scaling_factor = pre_act_scaling_factor_A * pre_act_scaling_factor_B    # :226
        self.act_scaling_factor = act_scaling_factor
        M, Kd, N = A.shape[-2], A.shape[-1], B.shape[-1]
        acc = K.bmm_i32(a_int.reshape(-1, M, Kd), b_int.reshape(-1, Kd, N), trans_b=False)
        out = K.int_to_carrier(acc, act_scaling_factor.reshape(-1))             # :228
        return out.reshape(*A.shape[:-1], N), act_scaling_factor


class QuantConv2d(nn.Conv2d):
    """quant_modules.py:231-330.  Only the form the models use (kernel == stride, no padding,
    groups == 1: a patch embedding) is implemented; it is an unfold + the tcgen05 GEMM."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, weight_bit=8, bias_bit=32, quant_mode="symmetric", per_channel=True,
                 weight_percentile=0):
        super(QuantConv2d, self).__init__(in_channels=in_channels, out_channels=out_channels,
                                          kernel_size=kernel_size, stride=stride, padding=padding,
                                          dilation=dilation, groups=groups, bias=bias)
        self.weight_bit = weight_bit
        self.quant_mode = quant_mode
        self.per_channel = per_channel
        self.weight_percentile = weight_percentile
        self.bias_bit = bias_bit
        self.quantize_bias = (False if bias_bit is None else True)
        self.register_buffer('conv_scaling_factor', torch.zeros(self.out_channels))
        self.register_buffer('weight_integer', torch.zeros_like(self.weight))
        self.register_buffer('bias_integer', torch.zeros_like(self.bias))
        self._w_cache = None

    def __repr__(self):
        s = super(QuantConv2d, self).__repr__()
        return "(" + s + " weight_bit={}, quant_mode={})".format(self.weight_bit, self.quant_mode)

    def fix(self):
        pass

    def unfix(self):
        pass

    def _quantized_weight(self):
        w = self.weight
        key = (w._version, w.data_ptr(), w.device)
        if self._w_cache is None or self._w_cache[0] != key:
            if not self.per_channel:
                raise Exception('For weight, we only support per_channel quantization.')
            with torch.no_grad():
                v = w.detach().reshape(w.shape[0], -1)                           # (c_in, kh, kw) order == unfold order
                self.min_val = v.min(axis=1).values
                self.max_val = v.max(axis=1).values
                s_w = symmetric_linear_quantization_params(self.weight_bit, self.min_val, self.max_val)
                w_q = K.quantize_f32(v, s_w, self.weight_bit, per_row=True)
            self._w_cache = (key, w_q, s_w)
            self.conv_scaling_factor = s_w
            self.weight_integer = w_q.to(torch.float32).reshape(w.shape)
        return self._w_cache[1], self._w_cache[2]

    def forward(self, x, pre_act_scaling_factor=None):
        if self.quant_mode == "asymmetric":
            raise NotImplementedError("unsupported quant mode: {}".format(self.quant_mode))
        elif self.quant_mode != "symmetric":
            raise ValueError("unknown quant mode: {}".format(self.quant_mode))
        p = self.kernel_size[0]
        if not (self.kernel_size == self.stride and self.kernel_size[0] == self.kernel_size[1]
                and self.padding == (0, 0) and self.dilation == (1, 1) and self.groups == 1):
            raise NotImplementedError("QuantConv2d: only kernel == stride patch embeddings are implemented "
                                      "(layers_quant.py:172-177)")
        w_q, s_w = self._quantized_weight()
        bias_scaling_factor = s_w * pre_act_scaling_factor                      # :321
        b_q = K.quantize_f32(self.bias.detach(), bias_scaling_factor, self.bias_bit, per_row=True,
                             out_dtype=torch.int32)                             # :322-323
        self.bias_integer = b_q.to(torch.float32)
        B, Cin, H, W = x.shape
        x_int = K.carrier_to_int(x.reshape(-1, 1), pre_act_scaling_factor.reshape(-1)[:1], torch.int8)   # :325-326
        patches = K.patchify_i8(x_int.reshape(B, Cin, H, W), p)
        out = K.gemm_i8(patches, w_q, bias=b_q, mode="carrier", scale=bias_scaling_factor)   # :329-330
        out = out.reshape(B, H // p, W // p, self.out_channels).permute(0, 3, 1, 2)
        return out, bias_scaling_factor.view(1, -1, 1, 1)


class IntLayerNorm(nn.LayerNorm):
    """I-LayerNorm, quant_modules.py:333-386."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super(IntLayerNorm, self).__init__(normalized_shape, eps, elementwise_affine)
        self.dim_sqrt = None
        self.register_buffer('norm_scaling_factor', torch.zeros(1))
        self.register_buffer('bias_integer', torch.zeros_like(self.bias))
        self._cache = None

    def fix(self):
        pass

    def unfix(self):
        pass

    def _static(self, C, device):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), device)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                n = torch.tensor(C, dtype=torch.float)
                self.dim_sqrt = torch.sqrt(n).to(device)                          # :355-356
                sf0 = self.dim_sqrt / 2 ** 30                                     # :374
                bias = self.bias.data.detach() / (self.weight.data.detach())      # :377
                bias_int = torch.floor(bias / sf0)                                # :378
                if float(bias_int.abs().max()) >= 2.0 ** 31:
                    raise OverflowError("IntLayerNorm: |beta/gamma| too large for the int32 bias path")
                out_sf = sf0 * self.weight.detach()                               # :383
            self._cache = (key, bias_int.to(torch.int32).contiguous(), out_sf.contiguous())
            self.bias_integer = bias_int
            self.norm_scaling_factor = out_sf
        return self._cache[1], self._cache[2]

    def forward(self, x, scaling_factor=None):
        if x.dim() != 3:
            raise NotImplementedError("IntLayerNorm normalises dim 2 of a [B, N, C] tensor (quant_modules.py:355,360)")
        bias_int, out_sf = self._static(x.shape[2], x.device)
        x_int = K.carrier_to_int(x, scaling_factor.reshape(-1), torch.int32)      # :359
        y = K.layernorm(x_int, bias_int)                                          # :360-382
        # :384-386.  The carrier is fp64: |y| reaches 2^30, beyond the 24-bit mantissa of fp32.  (The
        # reference's own module returns fp64 here when its input carrier is exact -- SURVEY.md 8c --
        # and a lossy fp32 product otherwise; the following QuantAct accepts either.)
        return K.int_to_carrier(y, out_sf, torch.float64), out_sf


def _host_scalar(t: torch.Tensor, cache: dict):
    """Value of a 1-element device tensor, cached on the tensor OBJECT and its version: frozen scales are the same
    tensor on every forward (QuantAct._frozen_scale), so the one device->host read happens once per operator, not per
    forward.  The cache holds a reference to the tensor, so its address cannot be recycled for another scale while the
    entry is alive; a scale produced by arithmetic (a new tensor every forward: calibration passes, `s * self.scale`)
    never hits and is read back each time."""
    if cache.get("tensor") is not t or cache.get("version") != t._version:
        cache["tensor"] = t
        cache["version"] = t._version
        cache["val"] = t.detach().reshape(-1)[:1].to(torch.float32).cpu()
    return cache["val"]


class IntGELU(nn.Module):
    """ShiftGELU, quant_modules.py:389-445."""

    def __init__(self, output_bit=8):
        super(IntGELU, self).__init__()
        self.output_bit = output_bit
        self.n = 23
        self.register_buffer('act_scaling_factor', torch.zeros(1))
        self._c = {}

    def fix(self):
        pass

    def unfix(self):
        pass

    def forward(self, x, scaling_factor=None):
        if self.output_bit != 8:
            raise NotImplementedError("IntGELU: output_bit != 8")
        s = scaling_factor.reshape(-1)
        if s.numel() != 1:
            raise NotImplementedError("IntGELU: scalar input scale expected")
        s_host = _host_scalar(scaling_factor, self._c)     # the producer's scale tensor object: stable when frozen
        x0 = int(torch.floor(-1.0 / (s_host * 1.702)))                            # :414, :427
        x_int = K.carrier_to_int(x, s, torch.int8)                                # :426
        y = K.shiftgelu(x_int, x0, n=self.n, out_dtype=torch.int32)               # :429-442
        sigmoid_scaling_factor = torch.tensor([1 / 2 ** (self.output_bit - 1)], dtype=torch.float32, device=x.device)
        out_sf = scaling_factor * sigmoid_scaling_factor                          # :443
        self.act_scaling_factor = out_sf
        return K.int_to_carrier(y, out_sf.reshape(-1)), out_sf


class IntSoftmax(nn.Module):
    """Shiftmax, quant_modules.py:448-497."""

    def __init__(self, output_bit=8):
        super(IntSoftmax, self).__init__()
        self.output_bit = output_bit
        self.n = 15
        self.register_buffer('act_scaling_factor', torch.zeros(1))
        self._c = {}

    def fix(self):
        pass

    def unfix(self):
        pass

    def forward(self, x, scaling_factor):
        if self.output_bit not in (8, 16):
            raise NotImplementedError("IntSoftmax: output_bit must be 8 or 16")
        s = scaling_factor.reshape(-1)
        if s.numel() != 1:
            raise NotImplementedError("IntSoftmax: scalar input scale expected")
        s_host = _host_scalar(scaling_factor, self._c)     # the producer's scale tensor object: stable when frozen
        x0 = int(torch.floor(-1.0 / s_host))                                      # :473
        x_int = K.carrier_to_int(x, s, torch.int32)                               # :484 (int32: Swin adds -100 masks)
        # below a scale of 2^-16 the reference's own result leaves [0, 2^(bits-1)) (clamped sum, :491-493): int32 storage
        p = K.shiftmax(x_int, x0, self.output_bit, n=self.n,
                       out_dtype=torch.int32 if x0 < -65536 else None)            # :485-493
        out_sf = torch.tensor([1 / 2 ** (self.output_bit - 1)], dtype=torch.float32, device=x.device)   # :494
        self.act_scaling_factor = out_sf
        return K.int_to_carrier(p, out_sf), out_sf
