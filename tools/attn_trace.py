"""Phase timeline of the pipelined attention kernel (CTA 0, clock64 stamps; needs a build with -DIVIT_ATTN_TRACE:
IVIT_NVCC_EXTRA=-DIVIT_ATTN_TRACE python i-vit_b200/csrc/build.py --force).

  python tools/attn_trace.py                  # plain pipelined kernel
  IVIT_ATTN_SWP=1 python tools/attn_trace.py  # tile-to-tile software-pipelined softmax warps

Softmax warps, per tile: events 0..7 (see AP_TR in ivit_attn_pipe.cu); control warp (16): 0 s_free seen, 1 S(t+1) issued,
2 p_ready seen, 3 o_free / V seen, 4 P V issued.  Prints, for a few warps, the cycles between consecutive events averaged
over tiles 8..39, and the tile period."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ivit_b200.kernels as K  # noqa: E402


def main():
    dev = torch.device("cuda")
    n_seq, n_tok, H, D = 256, 197, 12, 64
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randint(-128, 128, (n_seq * n_tok, 3 * H * D), dtype=torch.int8, device=dev, generator=g)
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
    torch.cuda.synchronize()
    dbg = torch.zeros(17 * 64 * 8, dtype=torch.int64, device=dev)
    os.environ["IVIT_ATTN_DBG_PTR"] = "%x" % dbg.data_ptr()
    run()
    torch.cuda.synchronize()
    del os.environ["IVIT_ATTN_DBG_PTR"]
    d = dbg.cpu().numpy().reshape(17, 64, 8)
    if not d.any():
        print("no stamps: library was not built with -DIVIT_ATTN_TRACE")
        return
    lo, hi = 8, 40
    print("checksum", int(out.to(torch.int64).sum().item()))
    for w in (0, 1, 2, 3, 4, 8, 12, 15, 16):
        ev = d[w, lo:hi].astype(np.int64)
        period = (d[w, hi, 0] - d[w, lo, 0]) / (hi - lo)
        nev = 5 if w == 16 else 8
        for par, name in ((0, "even tiles (m-tile 0)"), (1, "odd tiles (m-tile 1)")):
            e = ev[par::2]
            if w != 16 and not e[:, 2].any():
                seg = "inactive: " + " ".join("%d-%d:%6.0f" % (a, b, (e[:, b] - e[:, a]).mean()) for a, b in ((0, 1), (1, 6), (6, 7)) if e[:, a].all() and e[:, b].all())
            else:
                order = sorted(range(nev), key=lambda k: (e[:, k] - e[:, 0]).mean())
                seg = " ".join("%d>%d:%6.0f" % (order[i], order[i + 1], (e[:, order[i + 1]] - e[:, order[i]]).mean()) for i in range(nev - 1))
            print("warp %2d %-22s period/tile %6.0f | %s" % (w, name, period, seg))


if __name__ == "__main__":
    main()
