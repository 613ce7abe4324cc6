"""Micro-benchmark of the fused attention kernel at the DeiT-B bs=256 shape (CUDA events).

  python tools/attn_bench.py            # tcgen05 kernel when its preconditions hold
  IVIT_ATTN_TC=0 python tools/attn_bench.py   # mma.sync kernel
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ivit_b200.kernels as K  # noqa: E402


def main():
    dev = torch.device("cuda")
    n_seq, n_tok, H, D = int(os.environ.get("NSEQ", "256")), 197, int(os.environ.get("HEADS", "12")), 64
    qkv = torch.randint(-128, 128, (n_seq * n_tok, 3 * H * D), dtype=torch.int8, device=dev)
    s_attn = np.float32(0.031)
    acc_scale = np.float32(127 * s_attn / (D * 127 * 40))
    m_s, e_s = K.dyadic_host(np.array([acc_scale], np.float32), s_attn)
    x0 = int(np.floor(np.float32(-1.0) / s_attn))
    m_o, e_o = K.dyadic_host(np.array([2.0 ** -15 * 0.02], np.float32), np.float32(0.02 * 1.3))
    me_s, me_o = (int(m_s[0]), int(e_s[0])), (int(m_o[0]), int(e_o[0]))
    out = torch.empty((n_seq * n_tok, H * D), dtype=torch.int8, device=dev)
    run = lambda: K.attention_i8(qkv, n_seq, n_tok, H, D, me_s, x0, me_o, p_bits=16, out=out)
    for _ in range(3):
        run()
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("attention n_seq=%d n_tok=%d H=%d D=%d  %.4f ms  checksum %d" % (n_seq, n_tok, H, D, e0.elapsed_time(e1) / iters,
                                                                          int(out.to(torch.int64).sum().item())), flush=True)


if __name__ == "__main__":
    main()
