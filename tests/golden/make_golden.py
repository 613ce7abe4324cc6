"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN MODULES (this container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz / *.json

The reference has no tests, fixtures or known-answer vectors of its own (SURVEY.md
section 4), so parity is pinned by running its unmodified ``models/quantization_utils``
classes on seeded inputs (exact-carrier hooks of SURVEY.md section 8c where the fp32
carrier is not exact) and storing inputs + integer outputs:

  ops_kat.npz        per-operator known answers (every class of quantization_utils/__init__.py)
  deit_tiny_b2.npz   DeiT-tiny, batch 2: logits, sha256 of the integer tensor at every
                     operator boundary, full tensors for block 0 / tail, weight checksum
  swin_tiny_b1.npz / vit_large_b1.npz   the same digests for Swin-tiny (298 boundaries) and ViT-large (512), batch 1
  calib_<model>.json activation ranges (min,max per QuantAct) of the calibrated synthetic
                     models; weights are regenerated from the per-name seed (synth.py)

Nothing here is product code; nothing here runs on the GPU box.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refload  # noqa: E402
import ivit_b200  # noqa: E402
from ivit_b200.synth import synth_images, synth_parameters  # noqa: E402

torch.set_grad_enabled(False)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()


def set_range(qact, lo, hi):
    qact.min_val = torch.tensor(float(lo))
    qact.max_val = torch.tensor(float(hi))
    qact.running_stat = False


# ----------------------------------------------------------------------------- per-op KATs
def op_kats():
    m = refload.load()
    Q = m.quantization_utils.quant_modules
    U = m.quantization_utils.quant_utils
    rng = np.random.default_rng(1234)
    out = {}

    # --- batch_frexp (quant_utils.py:150-175) on fp64 ratios
    ratios = np.concatenate([
        rng.uniform(1e-6, 4.0, 40), -rng.uniform(1e-4, 2.0, 8),
        np.array([1.0, 0.5, 0.25, 2.0, 2.0 ** -40, 0.999999999999, 1234.5, 3e-12])])
    s_out = np.float32(0.0371)
    s_in = (ratios * np.float64(s_out)).astype(np.float32)
    mm, ee = U.batch_frexp(torch.from_numpy(s_in.astype(np.float64) / np.float64(s_out)))
    out["frexp_s_in"], out["frexp_s_out"] = s_in, np.float32(s_out)
    out["frexp_m"], out["frexp_e"] = mm.numpy().astype(np.int64), ee.numpy().astype(np.int64)

    # --- symmetric_linear_quantization_params (quant_utils.py:51-69)
    mins = rng.uniform(-3, 0.5, 12).astype(np.float32)
    maxs = rng.uniform(-0.5, 3, 12).astype(np.float32)
    mins[0], maxs[0] = 0.0, 0.0                      # -> eps clamp
    for b in (8, 16, 32):
        out["symscale_b%d" % b] = U.symmetric_linear_quantization_params(
            b, torch.from_numpy(mins), torch.from_numpy(maxs)).numpy()
    out["symscale_min"], out["symscale_max"] = mins, maxs

    # --- QuantAct input mode (quant_modules.py:194-196)
    for bits in (8, 16):
        x = (rng.standard_normal((2, 5, 24)) * 1.7).astype(np.float32)
        qa = Q.QuantAct(bits)
        set_range(qa, x.min() * 0.8, x.max() * 0.8)   # forces clamping
        y, sf = qa(torch.from_numpy(x))
        out["qin%d_x" % bits] = x
        out["qin%d_range" % bits] = np.array([qa.min_val.item(), qa.max_val.item()], np.float32)
        out["qin%d_sf" % bits] = sf.numpy().astype(np.float32)
        out["qin%d_q" % bits] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- QuantAct requant (quant_modules.py:197-206; fixedpoint_mul quant_utils.py:192-253)
    cases = []
    for ci, (bits, perch, resid, zmag, zdist) in enumerate([
            (8, False, False, 3000, "u"), (8, True, False, 200000, "u"), (16, True, False, 90000, "u"),
            (16, False, True, 30000, "u"), (16, True, True, 60000, "u"), (8, True, False, 2 ** 29, "ln"),
            (8, False, True, 127, "u"), (32, True, False, 50000, "u"), (8, True, False, 4000, "tie")]):
        rows, cols = 7, 40
        if zdist == "ln":
            z = rng.integers(-zmag, zmag, (rows, cols)).astype(np.int64)
        else:
            z = rng.integers(-zmag, zmag + 1, (rows, cols)).astype(np.int64)
        s_in = (rng.uniform(0.5, 2.0, cols if perch else 1) * 10.0 / zmag).astype(np.float32)
        if perch:
            s_in[::5] *= -1.0                         # negative per-channel scale (LN gamma < 0)
        qa = Q.QuantAct(bits)
        # 32-bit outputs are carried in fp32 by the reference (quant_utils.py:249): keep |q| < 2^24
        lim = 10.0 * 0.6 if bits != 32 else 8000.0
        set_range(qa, -lim, lim)
        if zdist == "tie":
            # power-of-two ratios -> m = 2^30 and exact ties z*m / 2^e = k + 0.5
            s_o = np.float32(lim / 127.0)
            s_in = (np.float32(s_o) * (2.0 ** -rng.integers(1, 6, cols))).astype(np.float32)
        pre = torch.from_numpy(z).double() * torch.from_numpy(s_in).double()
        kw = {}
        if resid:
            w = rng.integers(-20000, 20001, (rows, cols)).astype(np.int64)
            s_id = np.float32(0.00031)
            idt = torch.from_numpy(w).double() * float(s_id)
            y, sf = qa(pre, torch.from_numpy(s_in), idt, torch.tensor([s_id]))
            out["rq%d_w" % ci], out["rq%d_s_id" % ci] = w, s_id
        else:
            y, sf = qa(pre, torch.from_numpy(s_in))
        out["rq%d_z" % ci], out["rq%d_s_in" % ci] = z, s_in
        out["rq%d_sf" % ci] = sf.numpy().astype(np.float32)
        out["rq%d_q" % ci] = (y.double() / sf.double()).round().numpy().astype(np.int64)
        cases.append([ci, bits, int(perch), int(resid)])
    out["rq_cases"] = np.array(cases, np.int64)

    # --- QuantLinear (quant_modules.py:67-97)
    M, K, N = 9, 48, 20
    lin = Q.QuantLinear(K, N)
    lin.weight.copy_(torch.from_numpy((rng.standard_normal((N, K)) * 0.05).astype(np.float32)))
    lin.bias.copy_(torch.from_numpy((rng.standard_normal(N) * 0.2).astype(np.float32)))
    a = rng.integers(-128, 128, (M, K)).astype(np.int64)
    s_a = np.float32(0.0213)
    y, sf = lin(torch.from_numpy(a).float() * float(s_a), torch.tensor([s_a]))
    out["lin_w"], out["lin_b"] = lin.weight.numpy().copy(), lin.bias.numpy().copy()
    out["lin_a"], out["lin_s_a"] = a, s_a
    out["lin_wq"] = lin.weight_integer.numpy().astype(np.int64)
    out["lin_bq"] = lin.bias_integer.numpy().astype(np.int64)
    out["lin_sf"] = sf.numpy().astype(np.float32)
    out["lin_acc"] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- QuantConv2d (quant_modules.py:297-330): 4x4 / stride 4 patch embedding
    conv = Q.QuantConv2d(3, 8, kernel_size=4, stride=4)
    conv.weight.copy_(torch.from_numpy((rng.standard_normal((8, 3, 4, 4)) * 0.05).astype(np.float32)))
    conv.bias.copy_(torch.from_numpy((rng.standard_normal(8) * 0.2).astype(np.float32)))
    img = rng.integers(-128, 128, (2, 3, 8, 8)).astype(np.int64)
    s_i = np.float32(0.0171)
    y, sf = conv(torch.from_numpy(img).float() * float(s_i), torch.tensor([s_i]))
    out["conv_w"], out["conv_b"] = conv.weight.numpy().copy(), conv.bias.numpy().copy()
    out["conv_x"], out["conv_s"] = img, s_i
    out["conv_wq"] = conv.weight_integer.numpy().astype(np.int64)
    out["conv_bq"] = conv.bias_integer.numpy().astype(np.int64)
    out["conv_sf"] = sf.numpy().astype(np.float32).reshape(-1)
    out["conv_acc"] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- QuantMatMul (quant_modules.py:223-228): int8 x int8 and int16 x int8
    mmul = Q.QuantMatMul()
    A = rng.integers(-128, 128, (2, 3, 11, 16)).astype(np.int64)
    B = rng.integers(-128, 128, (2, 3, 16, 11)).astype(np.int64)
    sA, sB = np.float32(0.031), np.float32(0.027)
    y, sf = mmul(torch.from_numpy(A).float() * float(sA), torch.tensor([sA]),
                 torch.from_numpy(B).float() * float(sB), torch.tensor([sB]))
    out["mm_A"], out["mm_B"], out["mm_sA"], out["mm_sB"] = A, B, sA, sB
    out["mm_sf"] = sf.numpy().astype(np.float32)
    out["mm_acc"] = (y.double() / sf.double()).round().numpy().astype(np.int64)
    # rows of P sum to <= 2^15 as Shiftmax guarantees, so |acc| < 2^24 and the fp32 matmul is exact
    P = np.floor(rng.dirichlet(np.ones(11) * 0.3, (2, 3, 11)) * 32767).astype(np.int64)
    V = rng.integers(-128, 128, (2, 3, 11, 16)).astype(np.int64)
    sP = np.float32(2.0 ** -15)
    y, sf = mmul(torch.from_numpy(P).float() * float(sP), torch.tensor([sP]),
                 torch.from_numpy(V).float() * float(sB), torch.tensor([sB]))
    out["mm2_P"], out["mm2_V"] = P, V
    out["mm2_acc"] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- IntSoftmax (quant_modules.py:469-497), exact carrier
    for tag, bits, cols, s in [("sm16", 16, 197, 0.043), ("sm8", 8, 49, 0.0113),
                               ("sm16b", 16, 33, 0.25), ("sm8b", 8, 197, 0.0021)]:
        sm = Q.IntSoftmax(bits)
        q = rng.integers(-128, 128, (6, cols)).astype(np.int64)
        q[1] = -128
        q[2, :] = rng.integers(100, 128, cols)        # saturating row
        q[3, 0] = 127
        sc = torch.tensor([np.float32(s)])
        y, sf = sm(torch.from_numpy(q).double() * sc.double(), sc)
        out[tag + "_q"], out[tag + "_s"] = q, np.float32(s)
        out[tag + "_p"] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- IntGELU (quant_modules.py:410-445), exact carrier
    for tag, cols, s in [("gelu_a", 64, 0.0312), ("gelu_b", 768, 0.0521), ("gelu_c", 40, 0.0119)]:
        ge = Q.IntGELU()
        q = rng.integers(-128, 128, (6, cols)).astype(np.int64)
        q[1] = rng.integers(-128, -20, cols)          # all-negative row (k < 0 branch)
        q[2] = 0
        sc = torch.tensor([np.float32(s)])
        y, sf = ge(torch.from_numpy(q).double() * sc.double(), sc)
        out[tag + "_q"], out[tag + "_s"] = q, np.float32(s)
        out[tag + "_sf"] = sf.numpy().astype(np.float32)
        out[tag + "_o"] = (y.double() / sf.double()).round().numpy().astype(np.int64)

    # --- IntLayerNorm (quant_modules.py:353-386), exact carrier
    for tag, C, mag in [("ln_a", 192, 3000), ("ln_b", 768, 30000), ("ln_c", 96, 120), ("ln_d", 384, 32767)]:
        ln = Q.IntLayerNorm(C)
        g = rng.uniform(0.5, 1.5, C).astype(np.float32)
        g[::7] *= -1.0                                # negative gamma -> negative out scale
        ln.weight.copy_(torch.from_numpy(g))
        ln.bias.copy_(torch.from_numpy((rng.standard_normal(C) * 0.1).astype(np.float32)))
        q = rng.integers(-mag, mag + 1, (8, C)).astype(np.int64)
        q[1] = 5                                      # zero variance -> k = 64
        q[2] = 0
        q[2, 0] = mag                                 # one-hot row
        q[3] -= (q[3].sum() - C // 2) // C            # push the row mean toward a .5 tie
        q[3, 0] += (C // 2) - (q[3].sum() % C)        # exact tie: sum = k*C + C/2
        sc = torch.tensor([np.float32(0.00037)])
        x3 = (torch.from_numpy(q).double() * sc.double()).unsqueeze(0)
        y, sf = ln(x3, sc)
        out[tag + "_q"], out[tag + "_s"] = q, np.float32(0.00037)
        out[tag + "_g"], out[tag + "_beta"] = g, ln.bias.numpy().copy()
        out[tag + "_bq"] = ln.bias_integer.numpy().astype(np.int64)
        out[tag + "_sf"] = sf.detach().numpy().astype(np.float32)
        out[tag + "_o"] = (y[0].double() / sf.double()).round().numpy().astype(np.int64)
        out[tag + "_rowsum_mod"] = np.array([int(q[3].sum() % C)], np.int64)
    # --- round 2 additions.  A SEPARATE generator, so that every array above keeps its round-1 value.
    rng2 = np.random.default_rng(20262)
    # IntGELU at coarse input scales (x0 = floor(-1/(1.702 s)) in [-7, -1]: e^(-x_max) of an all-negative row reaches
    # 2^(23+184) and saturates the 2^31-1 clamp, quant_modules.py:434-438) and at a very fine one
    for tag, cols, s in [("gelu_x7", 96, 0.0840), ("gelu_x3", 64, 0.19), ("gelu_x2", 48, 0.29), ("gelu_x1", 80, 0.75),
                         ("gelu_fine", 64, 2.1e-5)]:
        ge = Q.IntGELU()
        q = rng2.integers(-128, 128, (7, cols)).astype(np.int64)
        q[1] = rng2.integers(-128, -20, cols)         # all-negative rows: -x_max > 0, k < 0
        q[2] = 0
        q[3] = -128
        q[4] = rng2.integers(-128, -100, cols)
        q[5, 0] = 127
        sc = torch.tensor([np.float32(s)])
        y, sf = ge(torch.from_numpy(q).double() * sc.double(), sc)
        out[tag + "_q"], out[tag + "_s"] = q, np.float32(s)
        out[tag + "_sf"] = sf.numpy().astype(np.float32)
        out[tag + "_o"] = (y.double() / sf.double()).round().numpy().astype(np.int64)
    # IntSoftmax at coarse / very fine scales (x0 = -1 ... -2^20)
    for tag, bits, cols, s in [("sm16_x1", 16, 50, 0.9), ("sm8_x2", 8, 49, 0.4), ("sm16_x5", 16, 197, 0.17),
                               ("sm8_fine", 8, 49, 1.1e-6), ("sm16_fine", 16, 64, 3.3e-5)]:
        sm = Q.IntSoftmax(bits)
        q = rng2.integers(-128, 128, (6, cols)).astype(np.int64)
        q[1] = -128
        q[2, :] = rng2.integers(100, 128, cols)
        q[3, 0] = 127
        sc = torch.tensor([np.float32(s)])
        y, sf = sm(torch.from_numpy(q).double() * sc.double(), sc)
        out[tag + "_q"], out[tag + "_s"] = q, np.float32(s)
        out[tag + "_p"] = (y.double() / sf.double()).round().numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "ops_kat.npz"), **out)
    print("ops_kat.npz:", len(out), "arrays")


# ----------------------------------------------------------------------------- whole model
def build_calibrated(factory_name, seed=0, calib_batch=8):
    m = refload.load()
    model = getattr(m, factory_name)(pretrained=False).eval()
    wsum = synth_parameters(model, seed)
    x = synth_images(calib_batch, seed)
    m.unfreeze_model(model)
    model(x)                                          # one unfrozen pass sets ranges (quant_train.py:266-311)
    m.freeze_model(model)                             # quant_train.py:325-326
    return model, wsum


def model_golden(factory_name, tag, batch, full_prefixes):
    model, wsum = build_calibrated(factory_name)
    calib = refload.calibration_table(model)
    with open(os.path.join(HERE, "calib_%s.json" % factory_name), "w") as f:
        json.dump({"model": factory_name, "seed": 0, "calib_batch": 8, "weights_sha256": wsum,
                   "ranges": calib}, f, indent=0, sort_keys=True)
    if batch is None:
        print("calib_%s.json written" % factory_name)
        return
    x = synth_images(batch, seed=7)
    y_lit = model(x).numpy().copy()                   # literal fp32 carrier (candidate 1)
    refload.add_exact_carrier_hooks(model)
    y, cap = refload.capture_integers(model, x)       # exact carrier (candidate 2 = the oracle of record)
    names = sorted(cap)
    out = {"logits": y.numpy().astype(np.float32), "logits_literal_fp32": y_lit.astype(np.float32),
           "names": np.array(names), "digests": np.array([digest(cap[n]) for n in names]),
           "shapes": np.array([json.dumps(list(cap[n].shape)) for n in names]),
           "weights_sha256": np.array(wsum), "seed_images": np.array(7), "batch": np.array(batch)}
    for n in names:
        if n in full_prefixes:
            a = cap[n]
            out["full/" + n] = a.astype(np.int32) if np.abs(a).max() < 2 ** 31 else a
    np.savez_compressed(os.path.join(HERE, "%s.npz" % tag), **out)
    print(tag, "boundaries:", len(names), "logit max|lit-exact| =", float(np.abs(y_lit - y.numpy()).max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "tiny", "calib"]
    if "ops" in which:
        op_kats()
    if "tiny" in which:
        model_golden("deit_tiny_patch16_224", "deit_tiny_b2", 2,
                     ["qact1", "blocks.0.attn.qact2", "blocks.0.qact4", "blocks.11.qact4", "qact2", "head"])
    if "calib" in which:
        for f in ("deit_small_patch16_224", "deit_base_patch16_224"):
            model_golden(f, None, None, [])
    if "large" in which:
        model_golden("vit_large_patch16_224", "vit_large_b1", 1, ["head"])
    if "swinb" in which:
        model_golden("swin_base_patch4_window7_224", "swin_base_b1", 1, ["head"])
    if "swin" in which:
        model_golden("swin_tiny_patch4_window7_224", "swin_tiny_b1", 1,
                     ["qact3", "head"])
