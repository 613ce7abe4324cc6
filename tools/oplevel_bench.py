"""Throughput of the operator-level drop-in path: the model graph stated on the reference's operator classes
(QuantLinear / QuantAct / IntLayerNorm / IntSoftmax / IntGELU / QuantMatMul, one C-ABI call or a few per operator),
i.e. what the reference's own model code gets when `models/quantization_utils` is swapped for this package --
next to the fused engine on the same frozen model.

  python tools/oplevel_bench.py [model] [batch]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ivit_b200.calib import build_synthetic  # noqa: E402
from ivit_b200.engine import Engine  # noqa: E402
from ivit_b200.pack import export_deit  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "deit_base_patch16_224"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    model = build_synthetic(name)
    x = torch.randn(B, 3, 224, 224, device="cuda")
    res = {}
    if name.startswith("deit"):
        eng = Engine(export_deit(model), "cuda")
        for _ in range(3):
            eng(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            eng(x)
        e1.record()
        torch.cuda.synchronize()
        res["engine_ms"] = e0.elapsed_time(e1) / 10
    model = model.cuda()
    with torch.no_grad():
        for _ in range(2):
            y = model(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = model(x)
        e1.record()
        torch.cuda.synchronize()
        res["oplevel_ms"] = e0.elapsed_time(e1) / 5
        res["oplevel_wall_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    print(name, "batch", B, " ".join("%s=%.2f" % kv for kv in res.items()),
          "| images/s operator-level %.0f" % (B / res["oplevel_ms"] * 1e3), flush=True)
    if name.startswith("deit"):
        print("  logits equal to the engine's:", bool(torch.equal(y.float(), eng(x).float())), flush=True)


if __name__ == "__main__":
    main()
