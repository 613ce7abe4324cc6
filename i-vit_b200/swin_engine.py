"""Fused executor for a frozen Swin parameter pack (``pack.export_swin``): integer tensors end to end, every launch
one of this package's sm_100a kernels -- no eager tensor op on the hot path.

Every QuantAct is fused into the kernel that produces its input, and the window glue of the reference graph
(``torch.roll`` / ``window_partition`` / ``window_reverse`` / the strided 2 x 2 merge + ``cat``, swin_quant.py:259-288,
337-341) is INDEX MATH inside a row kernel instead of copies:

* The residual stream of a block is kept in that block's WINDOW ORDER (tokens of a window contiguous, windows in
  partition order of the cyclically shifted grid).  LayerNorm / MLP / residual adds are row-wise, so they do not care;
  the qkv GEMM and the attention kernel see contiguous 49-token windows.
* Going from block b to block b+1 is one row permutation (reverse of b's shift + partition composed with b+1's), applied
  by norm1's row load (``ivit_layernorm_gather_i16_i8``, G = 1), which also writes the permuted int16 stream for the
  residual branch.  PatchMerging's gather + cat is the same kernel with four source rows per output row (G = 4).

    patch embedding      quantize+unfold (4x4) | tcgen05 GEMM (K = 48) + qact_before_norm | LayerNorm + patch_embed.qact
                         + qact1 (two 16-bit QuantActs) in one pass
    block                norm1+qact1 (gathered rows -> int8, + permuted int16 stream) | qkv GEMM (+attn.qact1) |
                         tcgen05 window attention: scores -> qact_attn1 -> qact2 (+ relative-position bias) -> mask ->
                         8-bit Shiftmax -> P V -> qact3 | proj GEMM (+attn.qact4, + residual QuantAct qact2 in the
                         epilogue) | norm2+qact3 | fc1 GEMM (+qact_gelu) | ShiftGELU table | fc2 GEMM (+mlp.qact2,
                         + residual QuantAct qact4)
    patch merging        norm+qact1 over 2x2-gathered rows | reduction GEMM (+qact2) | widen to the int16 stream
    head                 norm+qact2 | token average (RNE) + qact3 | head GEMM (fp32 logits)

Reference call order: swin_quant.py:539-564, 251-301, 121-169, 328-349.  Bit-identical to the CPU oracle
(oracle/model.py: swin_forward, itself pinned to the reference's digests at all 298 operator boundaries) --
tests/test_swin_gpu.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from ._lib import IvitError
from .pack import Pack


def _pair(v):
    return (int(v[0, 0]), int(v[0, 1]))


def window_order(R: int, ws: int, shift: int) -> np.ndarray:
    """Token index (raster, y * R + x) at every position of the window-ordered stream of a block: position
    ((wi * nW + wj) * ws + i) * ws + j holds the token that ``window_partition(torch.roll(x, (-shift, -shift)))`` puts
    there (swin_quant.py:259-271): rolled[y', x'] = x[(y' + shift) % R, (x' + shift) % R]."""
    nW = R // ws
    wi, wj, i, j = np.meshgrid(np.arange(nW), np.arange(nW), np.arange(ws), np.arange(ws), indexing="ij")
    y = (wi * ws + i + shift) % R
    x = (wj * ws + j + shift) % R
    return (y * R + x).reshape(-1).astype(np.int64)


def _inverse(order: np.ndarray) -> np.ndarray:
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    return inv


class SwinEngine:
    def __init__(self, pack: Pack, device="cuda", use_cuda_graph: bool = True):
        if pack.meta.get("arch") != "swin":
            raise NotImplementedError("SwinEngine: arch %r" % pack.meta.get("arch"))
        self.meta = m = dict(pack.meta)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ivit_b200.SwinEngine runs on a CUDA (sm_100a) device only")
        K.context(self.device)                      # fails loudly without the extension / a Blackwell GPU
        self.use_cuda_graph = use_cuda_graph
        dev_t = getattr(pack, "device_tensors", None)       # set by dist.broadcast_pack: already on the GPU
        if dev_t is not None and all(v.is_cuda for v in dev_t.values()):
            self.t = dict(dev_t)
        else:
            self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in pack.arrays.items()}
        t = self.t
        self.s = {k: _pair(v) for k, v in pack.arrays.items()
                  if (k.endswith(".me") or k.endswith(".me_res")) and v.shape[0] == 1}
        self.x0 = {k: int(v[0]) for k, v in pack.arrays.items() if k.endswith(".x0")}
        self.acc_bits = {}
        for k, v in pack.arrays.items():
            if k.endswith(".weight_integer"):
                name = k[:-len(".weight_integer")]
                bound = int(v.shape[1]) * 128 * 128 + int(np.abs(pack.arrays[name + ".bias_integer"].astype(np.int64)).max()) + 1
                self.acc_bits[name] = min(31, int(bound).bit_length())
        # static per-block tables: ShiftGELU table, requantised relative-position bias, mask bits, row permutations
        self.gelu_lut, self.bias_rq, self.mask_bits, self.mask_add, self.mask_i32 = {}, {}, {}, {}, {}
        self.rowmap, self.order, self.merge_map = {}, {}, {}
        R = m["grid"]
        order = np.arange(R * R, dtype=np.int64)            # the patch embedding leaves the stream in raster order
        for li, depth in enumerate(m["depths"]):
            ws = m["window"][li]
            for bi in range(depth):
                p = "layers.%d.blocks.%d." % (li, bi)
                self.gelu_lut[p] = K.shiftgelu_build_lut(self.x0[p + "mlp.act.x0"], t[p + "mlp.qact1.me"])
                # identity branch of attn.qact2 (swin_quant.py:149), static: RNE(bias * m_b / 2^e_b), int16
                b8 = t[p + "attn.bias_integer"]
                self.bias_rq[p] = K.requant(b8.reshape(-1, 1).to(torch.int32), t[p + "attn.qact2.me_res"], 16).reshape(b8.shape).contiguous()
                if p + "attn_mask" in pack.arrays:
                    s2 = np.float32(pack[p + "attn.qact2.scale"][0])
                    if not s2 <= np.float32(0.33):
                        raise ValueError("%sattn.qact2 scale %g: a masked score would not saturate Shiftmax (App. A.5)" % (p, s2))
                    add = int(np.rint(np.float64(-100.0) / np.float64(s2)))     # integer addend of a masked entry
                    m01 = pack[p + "attn_mask"].astype(np.uint64)
                    N = m01.shape[-1]
                    bits = (m01 << np.arange(N, dtype=np.uint64)[None, None, :]).sum(axis=2, dtype=np.uint64)
                    self.mask_bits[p] = torch.from_numpy(np.ascontiguousarray(bits).view(np.int64)).to(self.device)
                    self.mask_add[p] = add
                    self.mask_i32[p] = (t[p + "attn_mask"].to(torch.int32) * add).contiguous()   # general-kernel fallback
                new = window_order(R, ws, m["shift"][li][bi])
                self.order[p] = new
                if np.array_equal(new, order):
                    self.rowmap[p] = None
                else:
                    self.rowmap[p] = torch.from_numpy(_inverse(order)[new].astype(np.int32)).to(self.device)
                order = new
            if li + 1 < len(m["depths"]):
                d = "layers.%d.downsample." % li
                R2 = R // 2
                new = window_order(R2, m["window"][li + 1], 0)       # merged tokens in the next stage's first window order
                i2, j2 = new // R2, new % R2
                src = np.stack([(2 * i2) * R + 2 * j2, (2 * i2 + 1) * R + 2 * j2,            # x0, x1 (swin_quant.py:337-338)
                                (2 * i2) * R + 2 * j2 + 1, (2 * i2 + 1) * R + 2 * j2 + 1], 1)   # x2, x3 (:339-340)
                self.merge_map[d] = torch.from_numpy(_inverse(order)[src].reshape(-1).astype(np.int32)).to(self.device)
                self.order[d] = new
                order, R = new, R2
        self._plans = {}
        self.launches_per_forward = 0
        self.attention_fallbacks = 0

    def state_tensors(self):
        return self.t

    # ------------------------------------------------------------------ the launch sequence
    def _attention(self, p, qkv8, B_, nH, n_win_img):
        m, s = self.meta, self.s
        masked = p in self.mask_bits
        try:
            return K.window_attention_i8(qkv8, B_, nH, s[p + "attn.qact_attn1.me"], s[p + "attn.qact2.me"],
                                         self.x0[p + "attn.log_int_softmax.x0"], s[p + "attn.qact3.me"], self.bias_rq[p],
                                         mask_bits=self.mask_bits.get(p), n_win_img=n_win_img if masked else 0,
                                         mask_add=self.mask_add.get(p, 0))
        except IvitError as e:
            if "fast-form" not in str(e) and "only" not in str(e):
                raise
            self.attention_fallbacks += 1            # scales outside the tcgen05 kernel's domain: general kernel, same results
            N, D = m["window"][0] ** 2, qkv8.shape[1] // (3 * nH)
            return K.attention_i8(qkv8, B_, N, nH, D, s[p + "attn.qact_attn1.me"], self.x0[p + "attn.log_int_softmax.x0"],
                                  s[p + "attn.qact3.me"], p_bits=m["softmax_bits"], relbias=self.t[p + "attn.bias_integer"],
                                  me_s2=s[p + "attn.qact2.me"], me_b=s[p + "attn.qact2.me_res"], mask=self.mask_i32.get(p),
                                  n_win=(n_win_img if masked else 0))

    @staticmethod
    def _ln8(x, bias_int, me):
        """norm + QuantAct, int16 -> int8.  Narrow rows (C <= 384) take the gather kernel with the identity map: it packs up
        to eight rows into a warp (4 lanes per row at C = 96), where the DeiT-tuned kernel spends 16 lanes on a row."""
        rows, C = x.shape
        if C <= 384:
            return K.layernorm_gather(x, rows, C, 1, None, 1, 1, bias_int, me)
        return K.layernorm_i16_i8(x, bias_int, me)

    def _run(self, img: torch.Tensor, taps: dict = None):
        m, t, s = self.meta, self.t, self.s
        B = img.shape[0]
        P, G = m["patch"], m["grid"]
        C = m["embed_dim"]
        n = 0

        def tap(name, tensor, shape=None, order=None):
            """Record a boundary in the reference's layout: `order` (stream position -> token) un-permutes a stream tensor."""
            if taps is None:
                return
            x = tensor.reshape(shape) if shape is not None else tensor
            if order is not None:
                x = x[:, torch.from_numpy(_inverse(order)).to(x.device)]
            taps[name] = x.clone()

        def lin(name, a, me_key, bits, **kw):
            return K.gemm_i8(a, t[name + ".weight_integer"], bias=t[name + ".bias_integer"], mode="requant",
                             me=t[me_key + ".me"], bits=bits, acc_bits=self.acc_bits[name], **kw)

        patches = K.quantize_patchify(img, t["qact_input.scale"], P); n += 1            # swin_quant.py:540, layers_quant.py:190
        x8 = lin("patch_embed.proj", patches, "patch_embed.qact_before_norm", 8); n += 1  # :190 + :193
        tap("patch_embed.qact_before_norm", x8, (B, G * G, C))
        if taps is not None:                                                            # the two QuantActs one by one (diagnostic launches)
            x = K.layernorm(x8, t["patch_embed.norm.bias_integer"], t["patch_embed.qact.me"], bits=16)   # :194-195
            tap("patch_embed.qact", x, (B, G * G, C))
            tap("qact1", K.requant(x, t["qact1.me"], 16), (B, G * G, C))                # swin_quant.py:546
        # norm + patch_embed.qact (16 bit, per channel) + the model's qact1 (16 bit) in one pass over the int8 rows
        x = K.layernorm_i8_i16x2(x8, t["patch_embed.norm.bias_integer"], t["patch_embed.qact.me"], s["qact1.me"]); n += 1

        R = G
        for li, depth in enumerate(m["depths"]):
            nH, ws = m["num_heads"][li], m["window"][li]
            L = R * R
            n_win_img = (R // ws) ** 2
            B_ = B * n_win_img
            for bi in range(depth):
                p = "layers.%d.blocks.%d." % (li, bi)
                order = self.order[p]
                rmap = self.rowmap[p]
                if rmap is None:                                                        # stream already in this block's order
                    x1 = x
                    ln8 = self._ln8(x1, t[p + "norm1.bias_integer"], t[p + "qact1.me"]); n += 1   # :256-257
                else:                                                                   # :256-271 + :278-288 of the previous block
                    x1 = torch.empty_like(x)
                    ln8 = K.layernorm_gather(x, B * L, C, 1, rmap, L, L, t[p + "norm1.bias_integer"], t[p + "qact1.me"], xcopy=x1); n += 1
                tap(p + "qact1", ln8, (B, L, C), order)
                qkv8 = lin(p + "attn.qkv", ln8, p + "attn.qact1", 8); n += 1             # :128-129
                tap(p + "attn.qact1", qkv8, (B_, ws * ws, 3 * C))
                ao8 = self._attention(p, qkv8, B_, nH, n_win_img); n += 1                 # :135-164
                tap(p + "attn.qact3", ao8, (B_, ws * ws, C))
                if taps is not None:                                                     # attn.qact4 alone (diagnostic launch)
                    tap(p + "attn.qact4", lin(p + "attn.proj", ao8, p + "attn.qact4", 16), (B_, ws * ws, C))
                x2 = lin(p + "attn.proj", ao8, p + "attn.qact4", 16, two_stage=True, me2=s[p + "qact2.me"],
                         residual=x1, res_me=s[p + "qact2.me_res"]); n += 1               # :166-167 + residual QuantAct :293
                tap(p + "qact2", x2, (B, L, C), order)
                ln8 = self._ln8(x2, t[p + "norm2.bias_integer"], t[p + "qact3.me"]); n += 1   # :295-296
                tap(p + "qact3", ln8, (B, L, C), order)
                h8 = lin(p + "mlp.fc1", ln8, p + "mlp.qact_gelu", 8); n += 1             # layers_quant.py:145-146
                tap(p + "mlp.qact_gelu", h8, (B, L, -1), order)
                g8 = K.shiftgelu_lut(h8, self.gelu_lut[p]); n += 1                       # :147-148
                tap(p + "mlp.qact1", g8, (B, L, -1), order)
                x = lin(p + "mlp.fc2", g8, p + "mlp.qact2", 16, two_stage=True, me2=s[p + "qact4.me"],
                        residual=x2, res_me=s[p + "qact4.me_res"]); n += 1               # :150-151, swin_quant.py:299
                tap(p + "qact4", x, (B, L, C), order)
            if li + 1 < len(m["depths"]):                                               # PatchMerging :328-349
                d = "layers.%d.downsample." % li
                R //= 2
                L = R * R
                ln8 = K.layernorm_gather(x, B * L, 4 * C, 4, self.merge_map[d], L, 4 * L, t[d + "norm.bias_integer"],
                                         t[d + "qact1.me"]); n += 1                       # :337-345
                tap(d + "qact1", ln8, (B, L, 4 * C), self.order[d])
                C *= 2
                x8 = lin(d + "reduction", ln8, d + "qact2", 8); n += 1                   # :346-347
                tap(d + "qact2", x8, (B, L, C), self.order[d])
                x = K.widen_i8_i16(x8); n += 1                                           # the residual stream is carried as int16

        L = R * R
        ln8 = K.layernorm_i16_i8(x, t["norm.bias_integer"], t["qact2.me"]); n += 1        # :552-553
        tap("qact2", ln8, (B, L, C), order)
        z8 = K.avgpool_requant_i8(ln8, B, L, C, s["qact3.me"]); n += 1                   # :554-555 (order-independent)
        tap("qact3", z8, (B, C, 1))
        logits = K.gemm_i8(z8, t["head.weight_integer"], bias=t["head.bias_integer"], mode="carrier",
                           scale=t["head.out_scale"]); n += 1                            # :562
        if taps is None:
            self.launches_per_forward = n
        return logits

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """See ``_forward``; runs with the engine's device current (torch captures CUDA graphs on the current device)."""
        if torch.cuda.current_device() == (self.device.index if self.device.index is not None else torch.cuda.current_device()):
            return self._forward(images)
        with torch.cuda.device(self.device):
            return self._forward(images)

    def _forward(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [B, 3, H, W] on this engine's device -> fp32 logits [B, classes] (a view of an internal buffer,
        valid until the next call with the same batch size)."""
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
                self._run(static_in)                 # eager warm-up: function attributes, lazy initialisation, fallbacks
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):      # explicit stream: torch's default capture stream is global and may live on another device
                out = self._run(static_in)
            plan = self._plans[B] = {"in": static_in, "graph": g, "out": out}
        plan["in"].copy_(images, non_blocking=True)
        plan["graph"].replay()
        return plan["out"]

    __call__ = forward

    @torch.no_grad()
    def forward_taps(self, images: torch.Tensor) -> dict:
        """Eager forward that also returns the integer tensor at every fused-operator boundary (keyed by the reference
        module name of the LAST operator fused into that kernel), in the reference's token / window layout."""
        taps = {}
        taps["logits"] = self._run(images.contiguous().float(), taps).clone()
        return taps
