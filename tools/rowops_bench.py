"""Micro-benchmark of the HBM-bound row operators at the DeiT-B bs=256 shape (CUDA events; the operands of
one launch exceed L2 only for GELU, so a 256 MB scratch write flushes L2 between timed launches).

  python tools/rowops_bench.py            # LayerNorm int16->int8 [50432, 768], ShiftGELU LUT int8 [50432, 3072]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ivit_b200.kernels as K  # noqa: E402


def timed(run, flush, iters=10):
    for _ in range(3):
        run()
    tot = 0.0
    for _ in range(iters):
        flush.max()                                         # read-only pass over 256 MB: L2 holds clean lines of it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    dev = torch.device("cuda")
    M, C, Hd = int(os.environ.get("ROWS", "50432")), 768, 3072
    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn((M, C), generator=g) * 3000).clamp(-32768, 32767).to(torch.int16).to(dev)
    bias = torch.randint(-(1 << 20), 1 << 20, (C,), generator=g, dtype=torch.int32).to(dev)
    m, e = K.dyadic_host(np.linspace(2e-9, 4e-9, C).astype(np.float32), np.float32(0.02))
    me = K.dyadic_table(m, e, dev)
    out8 = torch.empty((M, C), dtype=torch.int8, device=dev)
    ms = timed(lambda: K.layernorm_i16_i8(x, bias, me, out=out8), flush)
    print("layernorm_i16_i8 [%d, %d]  %.4f ms  %.0f GB/s (3 B/element)  checksum %d" % (
        M, C, ms, 3.0 * M * C / ms / 1e6, int(out8.to(torch.int64).sum().item())), flush=True)
    q = torch.randint(-128, 128, (M, Hd), generator=g, dtype=torch.int8).to(dev)
    m1, e1 = K.dyadic_host(np.array([0.05 / 128], np.float32), np.float32(0.03))
    lut = K.shiftgelu_build_lut(-20, K.dyadic_table(m1, e1, dev))
    o = torch.empty_like(q)
    ms = timed(lambda: o.copy_(q), flush)
    print("torch copy int8  [%d, %d] %.4f ms  %.0f GB/s (same bytes as the GELU pass)" % (M, Hd, ms, 2.0 * M * Hd / ms / 1e6), flush=True)
    ms = timed(lambda: K.shiftgelu_lut(q, lut, out=o), flush)
    print("shiftgelu_lut    [%d, %d] %.4f ms  %.0f GB/s (2 B/element)  checksum %d" % (
        M, Hd, ms, 2.0 * M * Hd / ms / 1e6, int(o.to(torch.int64).sum().item())), flush=True)


if __name__ == "__main__":
    main()
