"""GPU whole-model parity: the fused engine and the operator-level drop-in path against the CPU
oracle (full tensors) and against the reference-generated golden digests, DeiT-tiny.  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle.model as OM
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<i8").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tiny():
    from ivit_b200.calib import build_synthetic
    from ivit_b200.engine import Engine
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_images
    model = build_synthetic("deit_tiny_patch16_224")
    pack = export_deit(model)
    gold = np.load(os.path.join(GOLDEN, "deit_tiny_b2.npz"))
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"]))
    cap = {}
    logits = OM.deit_forward(pack, x.numpy(), cap)
    eng = Engine(pack, "cuda")
    return dict(model=model, pack=pack, gold=gold, x=x, cap=cap, logits=logits, eng=eng)


def test_engine_matches_oracle_at_every_fused_boundary(tiny):
    taps = tiny["eng"].forward_taps(tiny["x"].cuda())
    cap = tiny["cap"]
    n = 0
    for name, t in taps.items():
        if name == "logits":
            continue
        got = t.cpu().numpy().astype(np.int64)
        want = cap[name]
        assert got.size == want.size, name
        bad = np.argwhere(got.reshape(-1) != want.reshape(-1))
        assert len(bad) == 0, "engine diverges from the oracle at %s: %d / %d elements, first flat index %d" % (
            name, len(bad), want.size, int(bad[0]))
        n += 1
    assert n >= 8 * 12 + 4
    assert np.array_equal(taps["logits"].cpu().numpy(), tiny["logits"]), "logits differ from the oracle"


def test_engine_matches_reference_digests(tiny):
    """Golden digests came from the reference's own modules: no oracle involved in this check."""
    gold = tiny["gold"]
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    shapes = dict(zip(gold["names"].tolist(), gold["shapes"].tolist()))
    taps = tiny["eng"].forward_taps(tiny["x"].cuda())
    n = 0
    for name, t in taps.items():
        if name in want and name != "qact_input":
            assert digest(t.cpu().numpy().astype(np.int64)) == want[name], "digest mismatch at %s" % name
            n += 1
    assert n >= 90
    err = np.abs(taps["logits"].cpu().numpy().astype(np.float64) - gold["logits"].astype(np.float64)).max()
    assert err <= 2e-6 * np.abs(gold["logits"]).max()        # see tests/test_model_oracle.py on the head's carrier noise
    assert (taps["logits"].cpu().numpy().argmax(1) == gold["logits"].argmax(1)).all()


def test_cuda_graph_replay_and_batch_sizes(tiny):
    from ivit_b200.synth import synth_images
    eng = tiny["eng"]
    xg = tiny["x"].cuda()
    a = eng(xg).clone()
    b = eng(xg).clone()                                      # second call replays the captured graph
    assert torch.equal(a, b)
    assert np.array_equal(a.cpu().numpy(), tiny["logits"])
    x5 = synth_images(5, seed=3)
    want = OM.deit_forward(tiny["pack"], x5.numpy())
    assert np.array_equal(eng(x5.cuda()).cpu().numpy(), want)
    assert np.array_equal(eng(x5[:1].cuda()).cpu().numpy(), want[:1])
    assert eng.launches_per_forward == 3 + 8 * 12 + 2        # fused stem: quantize+patchify, patch GEMM, embed; tail: norm (cls rows), head


def test_graph_bound_to_a_stable_input_buffer(tiny):
    """Feeding the same device buffer repeatedly binds a graph to its address (no copy into the engine's own input
    buffer); new contents of that buffer, and other buffers in between, must still give their own logits."""
    from ivit_b200.synth import synth_images
    eng = tiny["eng"]
    xa, xb = synth_images(3, seed=41), synth_images(3, seed=42)
    wa, wb = OM.deit_forward(tiny["pack"], xa.numpy()), OM.deit_forward(tiny["pack"], xb.numpy())
    buf = xa.cuda()
    for _ in range(4):                                       # copy path, then bound graph
        assert np.array_equal(eng(buf).cpu().numpy(), wa)
    buf.copy_(xb.cuda())                                     # same address, new images
    assert np.array_equal(eng(buf).cpu().numpy(), wb)
    other = xa.cuda()                                        # a different address: copy path again
    assert np.array_equal(eng(other).cpu().numpy(), wa)
    assert np.array_equal(eng(buf).cpu().numpy(), wb)


def test_uint8_input_path_matches_oracle(tiny):
    """Decoded uint8 pixels in: ToTensor + Normalize (utils/data_utils.py:90-91) happen inside the stem kernel; the logits
    must equal the oracle's on the images torch normalises in fp32 on the CPU."""
    eng = tiny["eng"]
    rng = np.random.default_rng(43)
    u = torch.from_numpy(rng.integers(0, 256, (3, 3, 224, 224)).astype(np.uint8))
    mean = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    x = u.to(torch.float32).div(255).sub(mean).div(std)
    want = OM.deit_forward(tiny["pack"], x.numpy())
    for _ in range(2):                                       # eager capture, then replay
        assert np.array_equal(eng(u.cuda()).cpu().numpy(), want)
    assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)  # the fp32 entry still agrees


def test_operator_level_path_matches_engine(tiny):
    """The drop-in operator classes (fp32 carrier in / out, one kernel per reference operator)
    give bit-identical logits to the fused engine."""
    from ivit_b200.quantization_utils import (IntGELU, IntLayerNorm, IntSoftmax, QuantAct, QuantConv2d, QuantLinear,
                                               QuantMatMul)
    model = tiny["model"].cuda()
    got, hooks = {}, []

    def mk(name):
        def hook(mod, inp, out):
            t, sf = out
            got[name] = (t.double() / sf.double()).round().to(torch.int64).cpu().numpy()
        return hook

    for name, mod in model.named_modules():
        if isinstance(mod, (QuantAct, QuantLinear, QuantConv2d, QuantMatMul, IntLayerNorm, IntSoftmax, IntGELU)):
            hooks.append(mod.register_forward_hook(mk(name)))
    with torch.no_grad():
        y = model(tiny["x"].cuda())
    for h in hooks:
        h.remove()
    model.cpu()
    cap = tiny["cap"]
    for name in cap:                                     # oracle (forward) order: report the FIRST divergence
        if name in got:
            bad = int((got[name].reshape(-1) != cap[name].reshape(-1)).sum())
            assert bad == 0, "operator-level path diverges from the oracle at %s (%d elements)" % (name, bad)
    assert len(set(cap) & set(got)) >= 250
    assert np.array_equal(y.cpu().numpy(), tiny["logits"])


def test_accelerate_replaces_the_forward_with_the_engine(tiny):
    """engine.accelerate(model): same logits as the operator-by-operator forward of the same model object."""
    import copy
    from ivit_b200.engine import accelerate
    model = copy.deepcopy(tiny["model"]).cuda()
    x = tiny["x"].cuda()
    with torch.no_grad():
        slow = model(x)
        accelerate(model)
        fast = model(x)
    assert hasattr(model, "_ivit_engine")
    assert np.array_equal(fast.cpu().numpy(), tiny["logits"])
    assert float((slow.float() - fast).abs().max()) <= 2e-6 * float(fast.abs().max())   # head carrier: fp32 product of the same integers


def test_operator_level_calibration_pass_runs(tiny):
    """One unfrozen forward (running_stat=True, quant_modules.py:170-189) through the operator
    classes on the GPU sets every executed QuantAct's range."""
    from ivit_b200 import deit
    from ivit_b200.model_utils import freeze_model, unfreeze_model
    from ivit_b200.synth import synth_images, synth_parameters
    m = deit.deit_tiny_patch16_224().eval()
    synth_parameters(m, 0)
    m = m.cuda()
    unfreeze_model(m)
    with torch.no_grad():
        m(synth_images(2, seed=0).cuda())
        freeze_model(m)
        y1 = m(synth_images(2, seed=5).cuda())
        y2 = m(synth_images(2, seed=5).cuda())
    assert torch.equal(y1, y2) and torch.isfinite(y1).all()
    assert float(torch.as_tensor(m.blocks[3].qact2.max_val)) > 0


def test_vit_large_fused_engine_matches_reference_digests():
    """ViT-large (C = 1024, 16 heads, 24 blocks) through the fused engine against the digests of the reference's own run
    (tests/golden/vit_large_b1.npz) -- no oracle involved."""
    from ivit_b200.calib import build_synthetic
    from ivit_b200.engine import Engine
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_images
    gold = np.load(os.path.join(GOLDEN, "vit_large_b1.npz"))
    want = dict(zip(gold["names"].tolist(), gold["digests"].tolist()))
    eng = Engine(export_deit(build_synthetic("vit_large_patch16_224")), "cuda")
    x = synth_images(int(gold["batch"]), seed=int(gold["seed_images"])).cuda()
    taps = eng.forward_taps(x)
    checked = 0
    for name, t in taps.items():
        if name in want and name != "qact_input":
            assert digest(t.cpu().numpy().astype(np.int64)) == want[name], "Engine (ViT-large) diverges from the reference at %s" % name
            checked += 1
    assert checked >= 24 * 8 + 2, checked
    y = eng(x).cpu().numpy()
    assert np.abs(y.astype(np.float64) - gold["logits"].astype(np.float64)).max() <= 2e-6 * np.abs(gold["logits"]).max()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(tiny):
    """one process driving two GPUs -- launches go to the context's device and stream whatever torch's current
    device is, ivit_create leaves the current device alone, and the per-kernel attributes are set per device."""
    from ivit_b200.engine import Engine
    assert torch.cuda.current_device() == 0
    eng1 = Engine(tiny["pack"], "cuda:1")
    assert torch.cuda.current_device() == 0                 # ivit_create restored it
    x = tiny["x"]
    y1 = eng1(x.to("cuda:1")).cpu().numpy()                  # current device is still 0
    y0 = tiny["eng"](x.to("cuda:0")).cpu().numpy()
    assert np.array_equal(y1, tiny["logits"]) and np.array_equal(y0, tiny["logits"])
    with torch.cuda.device(1):
        assert np.array_equal(tiny["eng"](x.to("cuda:0")).cpu().numpy(), tiny["logits"])   # engine on 0 while 1 is current
