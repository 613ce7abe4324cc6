"""Throughput of the fused Swin executor (BASELINE.json config 4: Swin-tiny INT8, bs=128, one B200) next to the
operator-by-operator drop-in path on the same frozen model.

  python tools/swin_bench.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ivit_b200.calib import build_synthetic  # noqa: E402
from ivit_b200.pack import export_swin  # noqa: E402
from ivit_b200.swin_engine import SwinEngine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    model = build_synthetic("swin_tiny_patch4_window7_224")
    eng = SwinEngine(export_swin(model), "cuda")
    x = torch.randn(B, 3, 224, 224, device="cuda")
    for _ in range(3):
        y = eng(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        y = eng(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("swin_tiny_patch4_window7_224 batch %d  fused engine %.2f ms  %.0f images/s" % (B, ms, B / ms * 1e3), flush=True)
    if os.environ.get("SWIN_OPLEVEL", "1") != "0":
        model = model.cuda()
        with torch.no_grad():
            model(x)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                z = model(x)
            e1.record()
            torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 3
        print("  operator-level path %.2f ms  %.0f images/s; logits equal: %s" % (ms2, B / ms2 * 1e3, bool(torch.equal(z.float(), y))), flush=True)


if __name__ == "__main__":
    main()
