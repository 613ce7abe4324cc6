"""Diagnose the first operator-level vs oracle divergence (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle.model as OM
from ivit_b200.calib import build_synthetic
from ivit_b200.pack import export_deit
from ivit_b200.synth import synth_images
from ivit_b200.quantization_utils import (IntGELU, IntLayerNorm, IntSoftmax, QuantAct, QuantConv2d, QuantLinear, QuantMatMul)
model = build_synthetic("deit_tiny_patch16_224")
pack = export_deit(model)
x = synth_images(2, seed=7)
cap = {}
OM.deit_forward(pack, x.numpy(), cap)
model = model.cuda()
got, sfs, ins = {}, {}, {}
def mk(name):
    def hook(mod, inp, out):
        t, sf = out
        got[name] = (t.double() / sf.double()).round().to(torch.int64).cpu().numpy()
        sfs[name] = sf.detach().float().cpu().numpy().reshape(-1)
        ins[name] = [i.detach().cpu() if torch.is_tensor(i) else i for i in inp]
    return hook
for name, mod in model.named_modules():
    if isinstance(mod, (QuantAct, QuantLinear, QuantConv2d, QuantMatMul, IntLayerNorm, IntSoftmax, IntGELU)):
        mod.register_forward_hook(mk(name))
with torch.no_grad():
    model(x.cuda())
for name in cap:
    if name not in got: continue
    g, w = got[name].reshape(-1), cap[name].reshape(-1)
    bad = np.flatnonzero(g != w)
    if len(bad):
        shp = cap[name].shape
        print("FIRST DIVERGENCE", name, "shape", shp, "nbad", len(bad), "max|want|", np.abs(w).max())
        idx = np.unravel_index(bad, shp)
        print(" distinct last-dim cols:", np.unique(idx[-1])[:20], " distinct rows:", np.unique(idx[-2])[:10] if len(shp) > 1 else "")
        for b in bad[:8]:
            print("   flat", b, "got", g[b], "want", w[b], "diff", g[b] - w[b])
        sf = sfs[name]
        cols = np.unique(idx[-1])
        print(" sf at bad cols", sf[cols[:8]] if sf.size > 1 else sf, " sf min/max", sf.min(), sf.max())
        if name.endswith("proj") or name.endswith("fc1") or name.endswith("qkv"):
            key = name
            print(" pack out_scale at cols", pack[key + ".out_scale"][cols[:8]], " bias_integer", pack[key + ".bias_integer"][cols[:8]])
            mod = dict(model.named_modules())[name]
            print(" module bias_integer", mod.bias_integer.cpu().numpy()[cols[:8]], " fc_scaling_factor", mod.fc_scaling_factor.cpu().numpy()[cols[:8]])
            print(" input sf", ins[name][1])
        break
else:
    print("NO DIVERGENCE")
