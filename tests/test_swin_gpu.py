"""Swin-tiny (windowed attention, relative position bias, shifted-window masks, patch merging) through the
operator-level drop-in classes on the GPU, against digests of the reference's own run (exact-carrier
hooks) at every operator boundary -- no oracle involved.  BASELINE.json config 4 (parity case)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()


def test_swin_tiny_operator_level_matches_reference_digests():
    from ivit_b200.calib import build_synthetic
    from ivit_b200.quantization_utils import (IntGELU, IntLayerNorm, IntSoftmax, QuantAct, QuantConv2d, QuantLinear,
                                               QuantMatMul)
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "swin_tiny_b1.npz"))
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    model = build_synthetic("swin_tiny_patch4_window7_224").cuda()
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).cuda()
    got = {}

    def mk(name):
        def hook(mod, inp, out):
            t, sf = out
            got[name] = digest((t.double() / sf.double()).round().to(torch.int64).cpu().numpy())
        return hook

    order = []
    for name, mod in model.named_modules():
        if isinstance(mod, (QuantAct, QuantLinear, QuantConv2d, QuantMatMul, IntLayerNorm, IntSoftmax, IntGELU)):
            mod.register_forward_hook(mk(name))
            order.append(name)
    with torch.no_grad():
        y = model(x)
    checked = 0
    for name in order:                               # registration order ~ forward order: report the first divergence
        if name in got and name in want:
            assert got[name] == want[name], "Swin operator-level path diverges from the reference at %s" % name
            checked += 1
    assert checked >= 290, checked
    err = np.abs(y.cpu().numpy().astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max()
    assert (y.cpu().numpy().argmax(1) == gold["logits"].argmax(1)).all()


def test_swin_fused_engine_matches_oracle_at_every_fused_boundary():
    """SwinEngine (integer tensors end to end, fused window attention with bias + mask, patch merging, token average)
    against the CPU oracle -- itself pinned to the reference's digests at all 298 boundaries -- at every boundary the
    engine materialises, for the golden batch and for a batch of 3 (graph replay included)."""
    import oracle.model as OM
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_swin
    from ivit_b200.swin_engine import SwinEngine
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "swin_tiny_b1.npz"))
    pack = export_swin(build_synthetic("swin_tiny_patch4_window7_224"))
    eng = SwinEngine(pack, "cuda")
    for batch, seed in [(int(gold["batch"]), int(gold["seed_images"])), (3, 19)]:
        x = synth_images(batch, seed=seed)
        cap = {}
        want = OM.swin_forward(pack, x.numpy(), cap)
        taps = eng.forward_taps(x.cuda())
        checked = 0
        for name in cap:                                  # oracle (forward) order: report the FIRST divergence
            if name in taps:
                got = taps[name].cpu().numpy().astype(np.int64).reshape(-1)
                bad = int((got != cap[name].reshape(-1)).sum())
                assert bad == 0, "SwinEngine diverges from the oracle at %s (%d of %d elements)" % (name, bad, got.size)
                checked += 1
        assert checked >= 12 * 9 + 3 * 2 + 5, checked
        assert np.array_equal(taps["logits"].cpu().numpy(), want)
        for _ in range(2):                                # captured graph, then replay
            assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)
    err = np.abs(want.astype(np.float64)).max()
    assert err > 0



def test_swin_base_fused_engine_matches_reference_digests():
    """Swin-base (C = 128 ... 1024, heads 4 / 8 / 16 / 32, merge LayerNorms up to 4 * 512 = 2048 channels) through the fused
    SwinEngine against the digests of the reference's own run (tests/golden/swin_base_b1.npz) -- no oracle involved."""
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_swin
    from ivit_b200.swin_engine import SwinEngine
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "swin_base_b1.npz"))
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    eng = SwinEngine(export_swin(build_synthetic("swin_base_patch4_window7_224")), "cuda")
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).cuda()
    taps = eng.forward_taps(x)
    checked = 0
    for name, t in taps.items():
        if name in want:
            assert digest(t.cpu().numpy().astype(np.int64)) == want[name], "SwinEngine (Swin-base) diverges from the reference at %s" % name
            checked += 1
    assert checked >= 24 * 9 + 3 * 2 + 4, checked
    y = eng(x).cpu().numpy()
    assert np.abs(y.astype(np.float64) - gold["logits"].astype(np.float64)).max() <= 2e-6 * np.abs(gold["logits"]).max()
    assert eng.attention_fallbacks == 0
