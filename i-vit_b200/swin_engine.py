"""Fused executor for a frozen Swin parameter pack (``pack.export_swin``): integer tensors end to end.

Same idea as ``engine.Engine`` for DeiT -- every QuantAct is fused into the kernel that produces its input, no fp32
carrier ever touches HBM -- with the kernels the DeiT path already has:

    patch embedding      quantize+unfold (4x4), tcgen05 GEMM (K = 48) + qact_before_norm, LayerNorm + 16-bit QuantAct
    block                LayerNorm+QuantAct (int16 -> int8) | window partition (cyclic shift) | qkv GEMM (+qact1) |
                         fused window attention: scores -> qact_attn1 -> qact2 with the relative-position bias as
                         identity -> shifted-window mask -> 8-bit Shiftmax -> P V -> qact3 (ivit_attention_i8) |
                         proj GEMM (+qact4, 16 bit) | window reverse | residual QuantAct | LayerNorm+QuantAct |
                         fc1 GEMM (+qact_gelu) | ShiftGELU table | fc2 GEMM (+qact2, residual QuantAct qact4)
    patch merging        2x2 gather | LayerNorm(4C)+QuantAct | reduction GEMM (+qact2)
    head                 LayerNorm+QuantAct | token average (RNE) | QuantAct | head GEMM (fp32 logits)

The window glue (roll, partition / reverse, the 2x2 gather) is plain tensor indexing on INTEGER tensors (int8 / int16,
a quarter / half of the bytes the reference moves there); everything arithmetic runs in the sm_100a kernels.  Reference
call order: swin_quant.py:539-564, 251-301, 121-169, 328-349.  Bit-identical to the CPU oracle (oracle/model.py:
swin_forward, itself pinned to the reference's digests at all 298 operator boundaries) -- tests/test_swin_gpu.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from .pack import Pack


def _pair(v):
    return (int(v[0, 0]), int(v[0, 1]))


class SwinEngine:
    def __init__(self, pack: Pack, device="cuda", use_cuda_graph: bool = True):
        if pack.meta.get("arch") != "swin":
            raise NotImplementedError("SwinEngine: arch %r" % pack.meta.get("arch"))
        self.meta = dict(pack.meta)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ivit_b200.SwinEngine runs on a CUDA (sm_100a) device only")
        K.context(self.device)                      # fails loudly without the extension / a Blackwell GPU
        self.use_cuda_graph = use_cuda_graph
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in pack.arrays.items()}
        self.s = {k: _pair(v) for k, v in pack.arrays.items()
                  if (k.endswith(".me") or k.endswith(".me_res")) and v.shape[0] == 1}
        self.x0 = {k: int(v[0]) for k, v in pack.arrays.items() if k.endswith(".x0")}
        self.gelu_lut, self.mask_add = {}, {}
        for li, depth in enumerate(self.meta["depths"]):
            for bi in range(depth):
                p = "layers.%d.blocks.%d." % (li, bi)
                self.gelu_lut[p] = K.shiftgelu_build_lut(self.x0[p + "mlp.act.x0"], self.t[p + "mlp.qact1.me"])
                if p + "attn_mask" in pack.arrays:
                    s2 = np.float32(pack[p + "attn.qact2.scale"][0])
                    if not s2 <= np.float32(0.33):
                        raise ValueError("%sattn.qact2 scale %g: a masked score would not saturate Shiftmax (App. A.5)" % (p, s2))
                    add = int(np.rint(np.float64(-100.0) / np.float64(s2)))     # integer addend of a masked entry
                    self.mask_add[p] = (self.t[p + "attn_mask"].to(torch.int32) * add).contiguous()
        self._plans = {}
        self.launches_per_forward = 0

    # ------------------------------------------------------------------ the launch sequence
    def _run(self, img: torch.Tensor, taps: dict = None):
        m, t, s = self.meta, self.t, self.s
        B = img.shape[0]
        P, G = m["patch"], m["grid"]
        C = m["embed_dim"]

        def tap(name, tensor, shape=None):
            if taps is not None:
                taps[name] = (tensor.reshape(shape) if shape is not None else tensor).clone()

        def lin(name, a, me_key, bits, **kw):
            return K.gemm_i8(a, t[name + ".weight_integer"], bias=t[name + ".bias_integer"], mode="requant",
                             me=t[me_key + ".me"], bits=bits, **kw)

        patches = K.quantize_patchify(img, t["qact_input.scale"], P)                    # swin_quant.py:540, layers_quant.py:190
        x8 = lin("patch_embed.proj", patches, "patch_embed.qact_before_norm", 8)         # :190 + :193
        tap("patch_embed.qact_before_norm", x8, (B, G * G, C))
        x = K.layernorm(x8, t["patch_embed.norm.bias_integer"], t["patch_embed.qact.me"], bits=16)   # :194-195
        tap("patch_embed.qact", x, (B, G * G, C))
        x = K.requant(x, t["qact1.me"], 16)                                             # swin_quant.py:546
        tap("qact1", x, (B, G * G, C))

        R = G
        for li, depth in enumerate(m["depths"]):
            nH, ws = m["num_heads"][li], m["window"][li]
            N, D, L = ws * ws, C // nH, R * R
            nWs = R // ws
            for bi in range(depth):
                p = "layers.%d.blocks.%d." % (li, bi)
                shift = m["shift"][li][bi]
                x1 = x                                                                  # int16 [B*L, C]
                ln8 = K.layernorm_i16_i8(x1, t[p + "norm1.bias_integer"], t[p + "qact1.me"])   # :256-257
                tap(p + "qact1", ln8, (B, L, C))
                g = ln8.view(B, R, R, C)
                if shift > 0:
                    g = torch.roll(g, shifts=(-shift, -shift), dims=(1, 2))             # :261-265
                xw = g.view(B, nWs, ws, nWs, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, C)   # :269-271
                B_ = B * nWs * nWs
                qkv8 = lin(p + "attn.qkv", xw, p + "attn.qact1", 8)                     # :128-129
                tap(p + "attn.qact1", qkv8, (B_, N, 3 * C))
                mask = self.mask_add.get(p)
                ao8 = K.attention_i8(qkv8, B_, N, nH, D, s[p + "attn.qact_attn1.me"], self.x0[p + "attn.log_int_softmax.x0"],
                                     s[p + "attn.qact3.me"], p_bits=m["softmax_bits"], relbias=t[p + "attn.bias_integer"],
                                     me_s2=s[p + "attn.qact2.me"], me_b=s[p + "attn.qact2.me_res"], mask=mask,
                                     n_win=(nWs * nWs if mask is not None else 0))      # :135-164
                tap(p + "attn.qact3", ao8, (B_, N, C))
                a16 = lin(p + "attn.proj", ao8, p + "attn.qact4", 16)                   # :166-167
                tap(p + "attn.qact4", a16, (B_, N, C))
                g = a16.view(B, nWs, nWs, ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, R, R, C)   # :278-281
                if shift > 0:
                    g = torch.roll(g, shifts=(shift, shift), dims=(1, 2))               # :284-288
                x2 = K.requant(g.reshape(B * L, C), t[p + "qact2.me"], 16, x1, t[p + "qact2.me_res"])   # :293
                tap(p + "qact2", x2, (B, L, C))
                ln8 = K.layernorm_i16_i8(x2, t[p + "norm2.bias_integer"], t[p + "qact3.me"])   # :295-296
                tap(p + "qact3", ln8, (B, L, C))
                h8 = lin(p + "mlp.fc1", ln8, p + "mlp.qact_gelu", 8)                    # layers_quant.py:145-146
                tap(p + "mlp.qact_gelu", h8, (B, L, -1))
                g8 = K.shiftgelu_lut(h8, self.gelu_lut[p])                              # :147-148
                tap(p + "mlp.qact1", g8, (B, L, -1))
                x = lin(p + "mlp.fc2", g8, p + "mlp.qact2", 16, two_stage=True, me2=s[p + "qact4.me"],
                        residual=x2, res_me=s[p + "qact4.me_res"])                      # :150-151, swin_quant.py:299
                tap(p + "qact4", x, (B, L, C))
            if li + 1 < len(m["depths"]):                                               # PatchMerging :328-349
                d = "layers.%d.downsample." % li
                g = x.view(B, R, R, C)
                g = torch.cat([g[:, 0::2, 0::2, :], g[:, 1::2, 0::2, :], g[:, 0::2, 1::2, :], g[:, 1::2, 1::2, :]], -1)   # :337-341
                R //= 2
                g = g.reshape(B * R * R, 4 * C)
                ln8 = K.layernorm(g, t[d + "norm.bias_integer"], t[d + "qact1.me"], bits=8)   # :344-345
                tap(d + "qact1", ln8, (B, R * R, 4 * C))
                C *= 2
                x8 = lin(d + "reduction", ln8, d + "qact2", 8)                          # :346-347
                tap(d + "qact2", x8, (B, R * R, C))
                x = x8.to(torch.int16)                                                  # the residual stream is carried as int16

        L = R * R
        ln8 = K.layernorm_i16_i8(x, t["norm.bias_integer"], t["qact2.me"])               # :552-553
        tap("qact2", ln8, (B, L, C))
        # token average, RNE(sum / L) (:554-555; exact integer reading, see oracle.avgpool_rne)
        ssum = ln8.view(B, L, C).sum(dim=1, dtype=torch.int32)
        qd = torch.div(ssum, L, rounding_mode="floor")
        rem = ssum - qd * L
        up = (2 * rem > L) | ((2 * rem == L) & ((qd & 1) == 1))
        pooled = (qd + up.to(torch.int32)).contiguous()
        z8 = K.requant(pooled, t["qact3.me"], 8)
        tap("qact3", z8, (B, C, 1))
        logits = K.gemm_i8(z8, t["head.weight_integer"], bias=t["head.bias_integer"], mode="carrier",
                           scale=t["head.out_scale"])                                   # :562
        return logits

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [B, 3, H, W] on this engine's device -> fp32 logits [B, classes]."""
        if images.device != self.device and not (images.is_cuda and self.device.index is None):
            raise RuntimeError("SwinEngine.forward: images on %s, engine on %s" % (images.device, self.device))
        images = images.contiguous().float()
        if not self.use_cuda_graph:
            return self._run(images)
        B = images.shape[0]
        plan = self._plans.get(B)
        if plan is None:
            static_in = torch.empty_like(images)
            static_in.copy_(images)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._run(static_in)                 # eager warm-up: function attributes, lazy initialisation
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._run(static_in)
            plan = self._plans[B] = {"in": static_in, "graph": g, "out": out}
        plan["in"].copy_(images, non_blocking=True)
        plan["graph"].replay()
        return plan["out"]

    __call__ = forward

    @torch.no_grad()
    def forward_taps(self, images: torch.Tensor) -> dict:
        """Eager forward that also returns the integer tensor at every fused-operator boundary (keyed by the reference
        module name of the LAST operator fused into that kernel)."""
        taps = {}
        taps["logits"] = self._run(images.contiguous().float(), taps).clone()
        return taps
