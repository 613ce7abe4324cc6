"""One eager fused Swin forward for ncu: python tools/profile_swin.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ivit_b200.calib import build_synthetic  # noqa: E402
from ivit_b200.pack import export_swin  # noqa: E402
from ivit_b200.swin_engine import SwinEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
eng = SwinEngine(export_swin(build_synthetic("swin_tiny_patch4_window7_224")), "cuda", use_cuda_graph=False)
x = torch.randn(B, 3, 224, 224, device="cuda")
eng(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
