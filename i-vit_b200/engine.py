"""Fused whole-model executor for a frozen DeiT/ViT parameter pack (the fast path).

One forward = a fixed sequence of sm_100a kernel launches on int8/int16 tensors that stay on
the device; every QuantAct of the reference graph is fused into the kernel that produces its
input (dyadic requant in the GEMM epilogue / LayerNorm / ShiftGELU / attention), the residual
adds ride in the proj / fc2 GEMM epilogues, scores and probabilities never leave the SM.
Per block (reference call order vit_quant.py:130-143, 59-88; layers_quant.py:144-153):

    norm1+qact1            ivit_layernorm  (int16 -> int8)
    qkv+attn.qact1         ivit_gemm_i8    (tcgen05, int8 out)
    matmul_1..attn.qact2   ivit_attention_i8
    proj+qact3+Block.qact2 ivit_gemm_i8    (two-stage requant + residual, int16 out)
    norm2+qact3            ivit_layernorm
    fc1+qact_gelu          ivit_gemm_i8    (int8 out)
    act+mlp.qact1          ivit_shiftgelu  (int8 -> int8)
    fc2+qact2+Block.qact4  ivit_gemm_i8    (two-stage requant + residual, int16 out)

The launch sequence is captured once into a CUDA graph per batch size and replayed.  Results
are bit-identical to the operator-level path (``quantization_utils``) and to the CPU oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from .pack import Pack, check_supported


class Engine:
    def __init__(self, pack: Pack, device="cuda", use_cuda_graph: bool = True):
        if pack.meta.get("arch") != "deit":
            raise NotImplementedError("Engine: arch %r" % pack.meta.get("arch"))
        check_supported(pack)
        self.meta = dict(pack.meta)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ivit_b200.Engine runs on a CUDA (sm_100a) device only")
        K.context(self.device)                      # fails loudly without the extension / a Blackwell GPU
        self.use_cuda_graph = use_cuda_graph
        dev_t = getattr(pack, "device_tensors", None)       # set by dist.broadcast_pack: already on the GPU
        if dev_t is not None and all(v.device == self.device or (v.is_cuda and self.device.index is None) for v in dev_t.values()):
            self.t = dict(dev_t)
        else:
            self.t = {k: torch.from_numpy(v).to(self.device) for k, v in pack.arrays.items()}
        self.s = {k: (int(v[0, 0]), int(v[0, 1])) for k, v in pack.arrays.items()
                  if (k.endswith(".me") or k.endswith(".me_res")) and v.shape[0] == 1}
        self.x0 = {k: int(v[0]) for k, v in pack.arrays.items() if k.endswith(".x0")}
        # IntGELU + mlp.qact1 over int8 is a function of (q, rowmax): one 64 KiB table per layer, built on the device
        self.gelu_lut = {i: K.shiftgelu_build_lut(self.x0["blocks.%d.mlp.act.x0" % i], self.t["blocks.%d.mlp.qact1.me" % i])
                         for i in range(self.meta["depth"])}
        # bound |acc + bias| < 2^acc_bits per linear layer (int8 operands: |acc| <= K * 128 * 128)
        self.acc_bits = {}
        for k, v in pack.arrays.items():
            if k.endswith(".weight_integer"):
                name = k[:-len(".weight_integer")]
                bound = int(v.shape[1]) * 128 * 128 + int(np.abs(pack.arrays[name + ".bias_integer"].astype(np.int64)).max()) + 1
                self.acc_bits[name] = min(31, int(bound).bit_length())
        self._plans = {}
        # eval transform of the reference (utils/data_utils.py:91, timm's IMAGENET_DEFAULT_MEAN / STD) for uint8 inputs
        self._norm = (torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32, device=self.device),
                      torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32, device=self.device))
        self._gemm_events = None
        self._cls_map = torch.zeros(1, dtype=torch.int32, device=self.device)      # x[:, 0] as a row map (vit_quant.py:272)
        self.launches_per_forward = 0

    # ------------------------------------------------------------------ parameter transport
    def state_tensors(self):
        """All device tensors of the frozen pack (for the one-time NCCL broadcast, dist.py)."""
        return self.t

    # ------------------------------------------------------------------ buffers
    def _buffers(self, B: int):
        m = self.meta
        C, N, Hd = m["embed_dim"], m["n_tok"], m["mlp_hidden"]
        M = B * N
        dev = self.device
        e = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=dev)
        return dict(
            img=e(B, m["in_chans"], m["img_size"], m["img_size"], dtype=torch.float32),
            img_q=e(B, m["in_chans"], m["img_size"], m["img_size"], dtype=torch.int8),
            img_u8=e(B, m["in_chans"], m["img_size"], m["img_size"], dtype=torch.uint8),
            patches=e(B * (N - 1), m["in_chans"] * m["patch"] ** 2, dtype=torch.int8),
            pe16=e(B * (N - 1), C, dtype=torch.int16),
            xa=e(M, C, dtype=torch.int16), xb=e(M, C, dtype=torch.int16),
            ln8=e(M, C, dtype=torch.int8), qkv8=e(M, 3 * C, dtype=torch.int8), ao8=e(M, C, dtype=torch.int8),
            h8=e(M, Hd, dtype=torch.int8), g8=e(M, Hd, dtype=torch.int8),
            cls16=e(B, C, dtype=torch.int16), cls8=e(B, C, dtype=torch.int8),
            logits=e(B, m["num_classes"], dtype=torch.float32))

    # ------------------------------------------------------------------ the launch sequence
    def _run(self, b, B: int, taps: dict = None, img: torch.Tensor = None):
        m, t, s = self.meta, self.t, self.s
        img = b["img"] if img is None else img
        u8 = img.dtype == torch.uint8                      # raw decoded image: ToTensor + Normalize fused into the stem kernel
        C, N, H, D = m["embed_dim"], m["n_tok"], m["num_heads"], m["head_dim"]
        n = 0

        def tap(name, tensor):
            if taps is not None:
                taps[name] = tensor.clone()

        def lin(name, a, out, me_key, bits, residual=None, stage2=None):
            kw = {}
            if stage2 is not None:                                   # per-channel QuantAct, then residual QuantAct
                kw = dict(two_stage=True, me2=s[stage2 + ".me"], residual=residual, res_me=s[stage2 + ".me_res"])
            ev = self._gemm_events
            if ev is not None:                                       # bench.py: CUDA events around every GEMM of an eager forward
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            K.gemm_i8(a, t[name + ".weight_integer"], bias=t[name + ".bias_integer"], mode="requant",
                      me=t[me_key + ".me"], bits=bits, out=out, acc_bits=self.acc_bits[name], **kw)
            if ev is not None:
                e1.record()
                w = t[name + ".weight_integer"]
                ev.append((name, a.shape[0], w.shape[0], w.shape[1], e0, e1))

        # input quantisation (vit_quant.py:257) -> patch unfold -> patch-embedding GEMM (+patch_embed.qact, 16 bit)
        if u8:
            K.quantize_patchify_u8(img, self._norm[0], self._norm[1], t["qact_input.scale"], m["patch"], out=b["patches"]); n += 1
        elif taps is None and m["patch"] % 4 == 0:
            K.quantize_patchify(img, t["qact_input.scale"], m["patch"], out=b["patches"]); n += 1   # one fused pass
        else:
            _quantize_into(img, t["qact_input.scale"], b["img_q"]); n += 1
            tap("qact_input", b["img_q"])
            K.call("ivit_patchify_i8", K.context(self.device), K.ptr(b["img_q"]), B, m["in_chans"], m["img_size"],
                   m["img_size"], m["patch"], K.ptr(b["patches"])); n += 1
        lin("patch_embed.proj", b["patches"], b["pe16"], "patch_embed.qact", 16); n += 1
        tap("patch_embed.qact", b["pe16"])
        # cls token + position embedding residual (vit_quant.py:259-265)
        e_ok = all(16 <= s[k][1] <= 62 for k in ("qact1.me", "qact1.me_res"))
        if e_ok and C % 8 == 0:
            K.embed_tokens_fast(b["pe16"], t["cls_token_integer"], t["pos_embed_integer"], B, N, C,
                                s["qact1.me"], s["qact1.me_res"], out=b["xa"]); n += 1
        else:
            K.embed_tokens(b["pe16"], t["cls_token_integer"], t["pos_embed_integer"], B, N, C,
                           s["qact1.me"], s["qact1.me_res"], 16, out=b["xa"]); n += 1
        tap("qact1", b["xa"])
        x, x2 = b["xa"], b["xb"]
        for i in range(m["depth"]):
            p = "blocks.%d." % i
            _layernorm_into(x, t[p + "norm1.bias_integer"], t[p + "qact1.me"], b["ln8"]); n += 1
            tap(p + "qact1", b["ln8"])
            lin(p + "attn.qkv", b["ln8"], b["qkv8"], p + "attn.qact1", 8); n += 1
            tap(p + "attn.qact1", b["qkv8"])
            K.attention_i8(b["qkv8"], B, N, H, D, s[p + "attn.qact_attn1.me"], self.x0[p + "attn.int_softmax.x0"],
                           s[p + "attn.qact2.me"], p_bits=m["softmax_bits"], out=b["ao8"]); n += 1
            tap(p + "attn.qact2", b["ao8"])
            lin(p + "attn.proj", b["ao8"], x2, p + "attn.qact3", 16, residual=x, stage2=p + "qact2"); n += 1
            tap(p + "qact2", x2)
            _layernorm_into(x2, t[p + "norm2.bias_integer"], t[p + "qact3.me"], b["ln8"]); n += 1
            tap(p + "qact3", b["ln8"])
            lin(p + "mlp.fc1", b["ln8"], b["h8"], p + "mlp.qact_gelu", 8); n += 1
            tap(p + "mlp.qact_gelu", b["h8"])
            K.shiftgelu_lut(b["h8"], self.gelu_lut[i], out=b["g8"]); n += 1
            tap(p + "mlp.qact1", b["g8"])
            lin(p + "mlp.fc2", b["g8"], x, p + "mlp.qact2", 16, residual=x2, stage2=p + "qact4"); n += 1
            tap(p + "qact4", x)
        # final norm on the cls rows only (LayerNorm is row-wise; vit_quant.py:271-273), then the head.  The x[:, 0] slice
        # is the row map of the gathering LayerNorm (one output row per image: token 0 of its N input rows)
        if C % 8 == 0 and C <= 2048:
            K.layernorm_gather(x, B, C, 1, self._cls_map, 1, N, t["norm.bias_integer"], t["qact2.me"], out=b["cls8"]); n += 1
        else:
            b["cls16"].copy_(x.view(B, N, C)[:, 0])
            _layernorm_into(b["cls16"], t["norm.bias_integer"], t["qact2.me"], b["cls8"]); n += 1
        tap("qact2", b["cls8"])
        K.gemm_i8(b["cls8"], t["head.weight_integer"], bias=t["head.bias_integer"], mode="carrier",
                  scale=t["head.out_scale"], out=b["logits"]); n += 1
        self.launches_per_forward = n
        return b["logits"]

    # ------------------------------------------------------------------ public API
    def _capture(self, b, B: int, img: torch.Tensor = None):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self._run(b, B, img=img)                 # eager warm-up: sets func attributes, builds nothing lazily later
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):      # explicit stream: torch's default capture stream is global and may live on another device
            self._run(b, B, img=img)
        return g

    def _plan(self, B: int):
        plan = self._plans.get(B)
        if plan is None:
            b = self._buffers(B)
            plan = {"buf": b, "graph": None, "graph_u8": None, "bound": {}, "seen": {}}
            if self.use_cuda_graph:
                plan["graph"] = self._capture(b, B)
            self._plans[B] = plan
        return plan

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """See ``_forward``.  Runs with the engine's device current (torch captures CUDA graphs on the CURRENT device: an
        engine on cuda:1 driven from a process whose current device is cuda:0 would otherwise capture nothing)."""
        if torch.cuda.current_device() == (self.device.index if self.device.index is not None else torch.cuda.current_device()):
            return self._forward(images)
        with torch.cuda.device(self.device):
            return self._forward(images)

    def _forward(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [B, 3, H, W] (already normalised, as the reference's models take them) or uint8 [B, 3, H, W]
        (decoded pixels: ToTensor + Normalize of the reference's eval transform are applied on the device, bit-identical
        to torchvision's fp32 arithmetic) on this engine's device -> fp32 logits [B, classes]
        (a view of an internal buffer, valid until the next call with the same batch size)."""
        if images.device != self.device and not (images.is_cuda and self.device.index is None):
            raise RuntimeError("Engine.forward: images on %s, engine on %s" % (images.device, self.device))
        B = images.shape[0]
        plan = self._plan(B)
        if images.dtype == torch.uint8:
            # decoded uint8 NCHW images: the reference's ToTensor + Normalize run inside the stem kernel (patch 16 models)
            if self.meta["patch"] != 16 or self.meta["in_chans"] > 4:
                raise NotImplementedError("uint8 input path: 16 x 16 patches, at most 4 channels")
            plan["buf"]["img_u8"].copy_(images, non_blocking=True)
            if plan["graph"] is None:
                self._run(plan["buf"], B, img=plan["buf"]["img_u8"])
            else:
                if plan["graph_u8"] is None:
                    plan["graph_u8"] = self._capture(plan["buf"], B, img=plan["buf"]["img_u8"])
                plan["graph_u8"].replay()
            return plan["buf"]["logits"]
        if plan["graph"] is None:
            plan["buf"]["img"].copy_(images, non_blocking=True)
            self._run(plan["buf"], B)
            return plan["buf"]["logits"]
        # A caller that feeds the same (contiguous fp32) buffer again gets a graph bound to that address: the 2 x 154 MB
        # device-to-device copy into the engine's own input buffer disappears (staging buffers of a serving loop,
        # bench.py).  First and second sight of an address still go through the copy; at most 4 addresses are bound.
        key = images.data_ptr()
        ok = images.dtype == torch.float32 and images.is_contiguous() and images.shape == plan["buf"]["img"].shape
        g = plan["bound"].get(key) if ok else None
        if g is None and ok:
            if len(plan["seen"]) > 256:                      # a caller that feeds fresh tensors every time: forget old addresses
                plan["seen"].clear()
            plan["seen"][key] = plan["seen"].get(key, 0) + 1
            if plan["seen"][key] >= 2 and len(plan["bound"]) < 4:
                g = plan["bound"][key] = self._capture(plan["buf"], B, img=images)
        if g is not None:
            g.replay()
        else:
            plan["buf"]["img"].copy_(images, non_blocking=True)
            plan["graph"].replay()
        return plan["buf"]["logits"]

    __call__ = forward

    @torch.no_grad()
    def time_gemms(self, images: torch.Tensor, forwards: int = 3):
        """Eager (un-graphed) forwards with a CUDA event pair around every tcgen05 GEMM launch of the step, on the
        launching stream: the kernels see the operands exactly as in the real forward (produced by the preceding
        kernel, residuals several kernels old).  Returns [(name, M, N, K, ms)] averaged over `forwards` runs."""
        B = images.shape[0]
        b = self._plan(B)["buf"]
        b["img"].copy_(images)
        self._run(b, B)                                   # warm-up
        acc = {}
        for _ in range(forwards):
            self._gemm_events = []
            self._run(b, B)
            torch.cuda.synchronize(self.device)
            for name, M, N, Kd, e0, e1 in self._gemm_events:
                a = acc.setdefault(name, [M, N, Kd, 0.0, 0])
                a[3] += e0.elapsed_time(e1)
                a[4] += 1
            self._gemm_events = None
        return [(k, v[0], v[1], v[2], v[3] / v[4]) for k, v in acc.items()]

    @torch.no_grad()
    def forward_taps(self, images: torch.Tensor) -> dict:
        """Eager forward that also returns the integer tensor at every fused-operator boundary
        (keyed by the reference module name of the LAST operator fused into that kernel)."""
        B = images.shape[0]
        b = self._buffers(B)
        b["img"].copy_(images)
        taps = {}
        logits = self._run(b, B, taps)
        taps["logits"] = logits.clone()
        return taps


# thin "write into a preallocated buffer" helpers (CUDA-graph friendly: no allocation in the hot loop)
def _quantize_into(x, scale, out):
    K.call("ivit_quantize_f32", K.context(x.device), K.ptr(x), x.numel(), K.ptr(scale), 1, 1, 8,
           K.TORCH2IVIT[out.dtype], K.ptr(out))


def _layernorm_into(x, bias_int, me, out):
    Cc = x.shape[-1]
    if x.dtype == torch.int16 and Cc % 8 == 0 and Cc <= 1024:
        K.layernorm_i16_i8(x, bias_int, me, out=out)          # vectorised hot-path form
    else:
        K.call("ivit_layernorm", K.context(x.device), K.ptr(x), K.TORCH2IVIT[x.dtype], x.numel() // Cc, Cc,
               K.ptr(bias_int), K.ptr(me), 8, K.TORCH2IVIT[out.dtype], K.ptr(out))


def _gelu_into(q, x0, me, out):
    cols = q.shape[-1]
    K.call("ivit_shiftgelu", K.context(q.device), K.ptr(q), K.TORCH2IVIT[q.dtype], q.numel() // cols, cols,
           int(x0), 23, K.ptr(me), 8, K.TORCH2IVIT[out.dtype], K.ptr(out))


def accelerate(model: torch.nn.Module, device="cuda"):
    """One-line adoption for code written against the reference's API: ``model = ivit_b200.engine.accelerate(model)``.

    ``model`` is a calibrated, frozen DeiT/ViT or Swin built on the reference's operator classes (the reference's own
    ``models/vit_quant.py`` / ``models/swin_quant.py`` object loaded through ``ivit_b200.dropin``, or this package's
    ``deit`` / ``swin`` graphs).  Its static integer parameters are exported once (``pack.export_deit`` /
    ``pack.export_swin``) and ``model.forward`` is replaced by the fused engine: same logits bit for bit as the
    operator-by-operator forward (tests/test_model_gpu.py, tests/test_zz_reference_graphs_gpu.py), ~30x its throughput
    (fp32 carriers never touch HBM).  Calling ``unfreeze_model`` / changing weights afterwards requires ``accelerate``
    again; the original forward is kept as ``model._ivit_forward_operator_level``."""
    from .pack import export_deit, export_swin
    for m in model.modules():
        if type(m).__name__ == "QuantAct" and getattr(m, "running_stat", False):
            raise RuntimeError("accelerate: the model is not frozen (QuantAct.running_stat is set); call freeze_model first")
    if hasattr(model, "layers") and hasattr(model, "patch_grid"):              # SwinTransformer (swin_quant.py:419)
        from .swin_engine import SwinEngine
        eng = SwinEngine(export_swin(model), device)
    elif hasattr(model, "blocks") and hasattr(model, "cls_token"):             # VisionTransformer (vit_quant.py:146)
        eng = Engine(export_deit(model), device)
    else:
        raise NotImplementedError("accelerate: %s is neither a VisionTransformer nor a SwinTransformer" % type(model).__name__)
    model._ivit_engine = eng
    model._ivit_forward_operator_level = model.forward

    def forward(x):
        return eng(x.to(eng.device, non_blocking=True)).clone()

    model.forward = forward
    return model
