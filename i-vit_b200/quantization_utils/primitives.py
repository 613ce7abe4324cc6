"""Mirror of the reference's ``models/quantization_utils/quant_utils.py`` forward halves
(inference only; the STE ``backward``s are training and out of scope).  Same names, same
argument meaning, same error behaviour; the arithmetic runs in the sm_100a kernels."""
from __future__ import annotations

import torch

from .. import kernels as K


def symmetric_linear_quantization_params(num_bits, min_val, max_val):
    """quant_utils.py:51-69 -- fp32: s = clamp(max(-min, max) / (2^(b-1)-1), eps).
    Host logic on (tiny) range tensors; evaluated once per operator after ``freeze_model``."""
    with torch.no_grad():
        n = 2 ** (num_bits - 1) - 1
        eps = torch.finfo(torch.float32).eps
        max_val = torch.max(-min_val, max_val)
        # tensor / tensor: a true IEEE division on every device.  (tensor / python-scalar is evaluated
        # as a multiplication by the reciprocal by torch's CUDA kernels, 1 ulp off the CPU result that
        # the frozen parameter pack and the oracle of record use.)
        scale = max_val / torch.tensor(float(n), dtype=max_val.dtype, device=max_val.device)
        scale = scale.clamp(min=eps)
    return scale


def linear_quantize(input, scale, zero_point, is_weight):
    """quant_utils.py:12-48 -- round(1/scale * input + zero_point); symmetric only (zero_point 0)."""
    if float(torch.as_tensor(zero_point).abs().max()) != 0.0:
        raise NotImplementedError("asymmetric quantization is not supported (reference: quant_modules.py:46,143)")
    if not is_weight and input.dim() > 4:
        raise NotImplementedError
    per_row = bool(is_weight) and scale.numel() > 1
    if not is_weight and scale.numel() > 1:
        raise NotImplementedError("per-channel activation quantisation of fp32 inputs is not used by the models")
    return K.quantize_f32(input, scale, 32, per_row=per_row, out_dtype=torch.int32)


class SymmetricQuantFunction:
    """quant_utils.py:72-96.  ``apply(x, k, specified_scale, is_weight)`` -> integer tensor
    (int8 / int16 / int32 storage instead of the reference's integer-valued fp32)."""

    @staticmethod
    def apply(x, k, specified_scale, is_weight):
        per_row = bool(is_weight) and specified_scale.numel() > 1
        if not is_weight and specified_scale.numel() > 1:
            raise NotImplementedError
        return K.quantize_f32(x, specified_scale, int(k), per_row=per_row)


class floor_ste:
    """quant_utils.py:122-133 (forward)."""

    @staticmethod
    def apply(x):
        return torch.floor(x)


class round_ste:
    """quant_utils.py:136-147 (forward)."""

    @staticmethod
    def apply(x):
        return torch.round(x)


def batch_frexp(inputs, max_bit=31):
    """quant_utils.py:150-175 for a ratio tensor ``inputs`` (fp64/fp32): returns (m, e) device
    tensors.  Unlike the reference there is no host round trip (ivit_dyadic kernel)."""
    if max_bit != 31:
        raise NotImplementedError("max_bit != 31")
    shape = inputs.shape
    one = torch.ones(1, dtype=torch.float32, device=inputs.device)
    t = K.dyadic_device(inputs.reshape(-1).float(), one)
    return t[:, 0].reshape(shape), t[:, 1].reshape(shape)


class fixedpoint_mul:
    """quant_utils.py:178-253 (forward): carrier in, integer-valued carrier out.

    ``apply(pre_act, pre_act_scaling_factor, bit_num, quant_mode, z_scaling_factor,
    identity=None, identity_scaling_factor=None)`` returns the integer tensor (fp32 values, as
    the reference does) -- QuantAct multiplies by the output scale."""

    @staticmethod
    def apply(pre_act, pre_act_scaling_factor, bit_num, quant_mode, z_scaling_factor,
              identity=None, identity_scaling_factor=None):
        q = fixedpoint_mul.integer(pre_act, pre_act_scaling_factor, bit_num, quant_mode, z_scaling_factor,
                                   identity, identity_scaling_factor)
        return q.to(torch.float32)

    @staticmethod
    def integer(pre_act, pre_act_scaling_factor, bit_num, quant_mode, z_scaling_factor,
                identity=None, identity_scaling_factor=None):
        if quant_mode != "symmetric":
            raise NotImplementedError("unsupported quant mode: {}".format(quant_mode))
        if pre_act.dim() not in (2, 3, 4):
            raise NotImplementedError
        s_in = pre_act_scaling_factor.reshape(-1)
        if pre_act.dim() == 4 and s_in.numel() > 1:
            raise NotImplementedError("per-channel scales on 4-D activations (dim 1) are not used by the models")
        if s_in.numel() not in (1, pre_act.shape[-1]):
            raise ValueError("scaling factor has %d entries for last dim %d" % (s_in.numel(), pre_act.shape[-1]))
        z = K.carrier_to_int_any(pre_act, s_in)              # integer shadow of the carrier when there is one
        me = K.dyadic_device(s_in, z_scaling_factor)
        w = me1 = None
        if identity is not None:
            s_id = identity_scaling_factor.reshape(-1)
            if identity.numel() > pre_act.numel() or pre_act.numel() % identity.numel() != 0:
                raise ValueError("identity shape %s does not broadcast to %s" % (tuple(identity.shape), tuple(pre_act.shape)))
            w = K.carrier_to_int_any(identity, s_id)
            me1 = K.dyadic_device(s_id, z_scaling_factor)
        return K.requant(z, me, int(bit_num), w, me1)
