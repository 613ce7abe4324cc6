"""One operator-level (drop-in classes) forward for ncu: python tools/profile_oplevel.py [model] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ivit_b200.calib import build_synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "deit_base_patch16_224"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = build_synthetic(name).cuda()
x = torch.randn(B, 3, 224, 224, device="cuda")
with torch.no_grad():
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
