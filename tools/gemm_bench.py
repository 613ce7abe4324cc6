"""Micro-benchmark of the tcgen05 GEMM on the DeiT-B bs=256 shapes (CUDA events, L2-warm operands).

  python tools/gemm_bench.py                 # current build, all shapes
  IVIT_GEMM_PAIR=0 python tools/gemm_bench.py
  IVIT_GEMM_DEBUG=1|2|3 python tools/gemm_bench.py     # 1: epilogue does no work, 2: producer loads nothing (diagnostics, wrong results)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ivit_b200.kernels as K  # noqa: E402

SHAPES = [("qkv", 50432, 2304, 768, 8, False), ("fc1", 50432, 3072, 768, 8, False),
          ("proj", 50432, 768, 768, 16, True), ("fc2", 50432, 768, 3072, 16, True),
          ("patch", 50176, 768, 768, 16, False)]          # patch embedding: plain 16-bit requant, no residual


def main():
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    iters = int(os.environ.get("ITERS", "20"))
    for name, M, N, Kd, bits, res in SHAPES:
        a = torch.randint(-128, 128, (M, Kd), dtype=torch.int8, device=dev)
        w = torch.randint(-128, 128, (N, Kd), dtype=torch.int8, device=dev)
        bias = torch.randint(-2 ** 15, 2 ** 15, (N,), dtype=torch.int32, device=dev)
        m = rng.integers(2 ** 30, 2 ** 31 - 1, N) | 1
        e = rng.integers(44, 48, N) if bits == 8 or res else rng.integers(33, 36, N)
        me = K.dyadic_table(m, e, dev)
        out = torch.empty((M, N), dtype=torch.int8 if bits == 8 else torch.int16, device=dev)
        kw = {}
        if res:
            r = torch.randint(-30000, 30000, (M, N), dtype=torch.int16, device=dev)
            kw = dict(two_stage=True, me2=(2 ** 30 + 12345, 33), residual=r, res_me=(2 ** 30 + 777, 31))
        run = lambda: K.gemm_i8(a, w, bias=bias, mode="requant", me=me, bits=bits, out=out, acc_bits=24, **kw)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print("%-5s M=%d N=%d K=%d  %.4f ms  %.0f TOPS" % (name, M, N, Kd, ms, 2.0 * M * N * Kd / (ms * 1e-3) / 1e12), flush=True)


if __name__ == "__main__":
    main()
