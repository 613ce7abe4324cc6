import os, sys
sys.path.insert(0, os.getcwd())
import torch
import ivit_b200.kernels as K
for (b, M, Kd, N, tb) in [(768, 197, 64, 197, False), (768, 197, 197, 64, False), (3072, 197, 64, 197, False)]:
    a = torch.randint(-30000, 30000, (b, M, Kd), dtype=torch.int16, device="cuda")
    w = torch.randint(-128, 128, (b, Kd, N), dtype=torch.int8, device="cuda")
    for _ in range(2): c = K.bmm_i32(a, w, trans_b=tb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): c = K.bmm_i32(a, w, trans_b=tb)
    e1.record(); torch.cuda.synchronize()
    ref = torch.bmm(a[:4].double(), w[:4].double()).to(torch.int32)
    print("bmm", b, M, Kd, N, "%.3f ms" % (e0.elapsed_time(e1) / 5), "ok" if torch.equal(c[:4], ref) else "MISMATCH", flush=True)
