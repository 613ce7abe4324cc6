"""Whole-model pinning (CPU): frozen parameter pack + oracle forward reproduce, at EVERY operator
boundary, the integers obtained by running the reference's own DeiT-tiny with exact-carrier hooks
(tests/golden/deit_tiny_b2.npz, made by tests/golden/make_golden.py), and its logits bit for bit."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

import oracle.model as OM
from conftest import GOLDEN


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tiny():
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_images
    model = build_synthetic("deit_tiny_patch16_224")
    pack = export_deit(model)
    gold = np.load(os.path.join(GOLDEN, "deit_tiny_b2.npz"))
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).numpy()
    return model, pack, gold, x


def test_oracle_reproduces_reference_at_every_boundary(tiny):
    model, pack, gold, x = tiny
    cap = {}
    logits = OM.deit_forward(pack, x, cap)
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    checked = 0
    for name, arr in cap.items():
        assert name in want, "oracle boundary %s is not a reference module" % name
        assert digest(arr) == want[name], "first divergence from the reference at %s" % name
        checked += 1
    # every reference boundary that was actually executed is covered
    executed = {n for n in want if "qact_softmax" not in n and n != "act_out"}
    assert executed <= set(cap), "boundaries not restated by the oracle: %s" % sorted(executed - set(cap))[:5]
    assert checked >= 250
    # The head's integer accumulator matched above (boundary "head").  Its fp32 logits cannot be
    # bit-identical to the reference's: the reference feeds x / s (NOT rounded, quant_modules.py:94)
    # into an fp32 F.linear with no QuantAct behind it to absorb the carrier noise, so its logits
    # sit within a few ulp of fp32(acc) * scale, which is what the oracle and the engine compute.
    # (noise ~1e-7 relative to the accumulation magnitude, not to the possibly tiny logit)
    err = np.abs(logits.astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max(), "logits differ from the reference by %g" % err
    for key in gold.files:
        if key.startswith("full/"):
            assert np.array_equal(cap[key[5:]].reshape(gold[key].shape), gold[key])
    # the literal fp32-carrier reference differs only by carrier noise
    lit = gold["logits_literal_fp32"]
    assert np.abs(lit - logits).max() < 0.1 and (lit.argmax(1) == logits.argmax(1)).all()


def test_pack_roundtrip_and_domains(tiny, tmp_path):
    from ivit_b200.pack import Pack, check_supported
    _, pack, _, _ = tiny
    check_supported(pack)
    f = str(tmp_path / "p.npz")
    pack.save(f)
    p2 = Pack.load(f)
    assert p2.meta == pack.meta and set(p2.arrays) == set(pack.arrays)
    for k in pack.arrays:
        assert np.array_equal(p2[k], pack[k]) and p2[k].dtype == pack[k].dtype
    # the fast GEMM epilogue path needs 32 <= e <= 62 for the linear outputs of realistic scales
    for k, v in pack.arrays.items():
        if k.endswith(".me"):
            assert (np.abs(v[:, 0].astype(np.int64)) >= 2 ** 30).all() and (v[:, 1] >= -1).all() and (v[:, 1] <= 63).all()


def test_pack_from_reference_model_is_identical(tiny):
    sys.path.insert(0, GOLDEN)
    import refload
    if not refload.have_reference():
        pytest.skip("reference checkout not present (GPU box)")
    from ivit_b200.calib import apply_calibration, load_calibration
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_parameters
    _, pack, _, _ = tiny
    m = refload.load()
    with torch.no_grad():
        ref = m.deit_tiny_patch16_224(pretrained=False).eval()
        cal = load_calibration("deit_tiny_patch16_224")
        assert synth_parameters(ref, 0) == cal["weights_sha256"]
        apply_calibration(ref, cal["ranges"])
        pref = export_deit(ref)                      # the exporter reads the REFERENCE's model object
    assert pref.meta == pack.meta
    for k in pack.arrays:
        assert np.array_equal(pref[k], pack[k]), k


# ------------------------------------------------------------------------------------------------ Swin
@pytest.fixture(scope="module")
def swin():
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_swin
    from ivit_b200.synth import synth_images
    model = build_synthetic("swin_tiny_patch4_window7_224")
    pack = export_swin(model)
    gold = np.load(os.path.join(GOLDEN, "swin_tiny_b1.npz"))
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).numpy()
    return model, pack, gold, x


def test_swin_oracle_reproduces_reference_at_every_boundary(swin):
    """Swin-tiny (BASELINE.json config 4): frozen pack + oracle forward -- windows, cyclic shift + mask, relative-position
    bias as a QuantAct identity, patch merging, token average -- against the digests of the reference's own run with
    exact-carrier hooks at all 298 operator boundaries (tests/golden/swin_tiny_b1.npz)."""
    model, pack, gold, x = swin
    cap = {}
    logits = OM.swin_forward(pack, x, cap)
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    for name, arr in cap.items():
        assert name in want, "oracle boundary %s is not a reference module" % name
        assert digest(arr) == want[name], "first divergence from the reference at %s" % name
    assert set(want) <= set(cap), "boundaries not restated by the oracle: %s" % sorted(set(want) - set(cap))[:5]
    assert len(cap) >= 298
    err = np.abs(logits.astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max()
    for key in gold.files:
        if key.startswith("full/"):
            assert np.array_equal(cap[key[5:]].reshape(gold[key].shape), gold[key])
    # shifted blocks really exercised the mask path
    assert any(k.endswith("attn_mask") for k in pack.arrays)


def test_swin_pack_roundtrip_and_reference_model(swin, tmp_path):
    from ivit_b200.pack import Pack, export_swin
    _, pack, _, _ = swin
    f = str(tmp_path / "s.npz")
    pack.save(f)
    p2 = Pack.load(f)
    assert p2.meta == pack.meta and all(np.array_equal(p2[k], pack[k]) for k in pack.arrays)
    sys.path.insert(0, GOLDEN)
    import refload
    if not refload.have_reference():
        pytest.skip("reference checkout not present (GPU box)")
    from ivit_b200.calib import apply_calibration, load_calibration
    from ivit_b200.synth import synth_parameters
    m = refload.load()
    with torch.no_grad():
        ref = m.swin_tiny_patch4_window7_224(pretrained=False).eval()
        cal = load_calibration("swin_tiny_patch4_window7_224")
        assert synth_parameters(ref, 0) == cal["weights_sha256"]
        apply_calibration(ref, cal["ranges"])
        pref = export_swin(ref)                      # the exporter reads the REFERENCE's model object
    assert pref.meta == pack.meta
    for k in pack.arrays:
        assert np.array_equal(pref[k], pack[k]), k


def test_vit_large_oracle_reproduces_reference_at_every_boundary():
    """ViT-large (24 blocks, C = 1024, 16 heads; vit_quant.py:365-381): pack + oracle forward against the digests of the
    reference's own run at all 512 operator boundaries (tests/golden/vit_large_b1.npz)."""
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "vit_large_b1.npz"))
    pack = export_deit(build_synthetic("vit_large_patch16_224"))
    assert pack.meta["depth"] == 24 and pack.meta["embed_dim"] == 1024
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).numpy()
    cap = {}
    logits = OM.deit_forward(pack, x, cap)
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    for name, arr in cap.items():
        assert digest(arr) == want[name], "first divergence from the reference at %s" % name
    executed = {n for n in want if "qact_softmax" not in n and n != "act_out"}
    assert executed <= set(cap) and len(cap) >= 500
    err = np.abs(logits.astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max()


def test_swin_base_oracle_reproduces_reference_at_every_boundary():
    """Swin-base (depths 2/2/18/2, heads 4/8/16/32, C = 128..1024; swin_quant.py:609-627): pack + oracle forward against
    the digests of the reference's own run at all 574 operator boundaries (tests/golden/swin_base_b1.npz)."""
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_swin
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "swin_base_b1.npz"))
    pack = export_swin(build_synthetic("swin_base_patch4_window7_224"))
    assert pack.meta["depths"] == [2, 2, 18, 2] and pack.meta["num_heads"] == [4, 8, 16, 32]
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).numpy()
    cap = {}
    logits = OM.swin_forward(pack, x, cap)
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    for name, arr in cap.items():
        assert digest(arr) == want[name], "first divergence from the reference at %s" % name
    assert set(want) <= set(cap) and len(cap) >= 570
    err = np.abs(logits.astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max()

