"""Run N eager (no CUDA graph) forwards of the fused engine -- the command ncu wraps:

  ncu --metrics gpu__time_duration.sum --clock-control none -s <L> -c <L> --csv --log-file gpurun_out/launches.csv \
      python tools/profile_forward.py --batch 256 --forwards 2        (L = launches per forward, printed below)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ivit_b200.calib import build_synthetic  # noqa: E402
from ivit_b200.engine import Engine  # noqa: E402
from ivit_b200.pack import export_deit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="deit_base_patch16_224")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--forwards", type=int, default=2)
a = ap.parse_args()
eng = Engine(export_deit(build_synthetic(a.model)), "cuda", use_cuda_graph=False)
x = torch.randn(a.batch, 3, 224, 224, device="cuda")
for _ in range(a.forwards):
    eng(x)
torch.cuda.synchronize()
print("launches_per_forward", eng.launches_per_forward)
