"""the frozen pack straight from a saved QAT checkpoint (``state_dict`` keys of the reference), and the
TVM converter's on-disk parameters from a pack.  CPU only; the checkpoints are produced by the UNMODIFIED reference
(tests/golden/refload.py), so these tests run where /root/reference (or the staged baseline/_ref) exists."""
import ast
import io
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import refload  # noqa: E402

pytestmark = pytest.mark.skipif(not refload.have_reference(), reason="reference checkout not present")


def _frozen_reference_checkpoint(name, batch=1):
    """Reference model -> synthetic weights -> golden calibration -> frozen forward -> torch.save(state_dict) bytes."""
    from ivit_b200.calib import apply_calibration, load_calibration
    from ivit_b200.synth import synth_images, synth_parameters
    m = refload.load()
    model = getattr(m, name)(pretrained=False).eval()
    cal = load_calibration(name)
    assert synth_parameters(model, cal["seed"]) == cal["weights_sha256"]
    apply_calibration(model, cal["ranges"])
    m.freeze_model(model)
    with torch.no_grad():
        model(synth_images(batch, seed=5))                 # fills weight_integer / bias_integer / *scaling_factor buffers
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)                    # quant_train.py:261
    buf.seek(0)
    return model, buf


@pytest.mark.parametrize("name", ["deit_tiny_patch16_224", "swin_tiny_patch4_window7_224"])
def test_pack_from_a_saved_state_dict_equals_the_pack_of_the_live_model(name):
    from ivit_b200.pack import Pack, export_deit, export_swin
    model, buf = _frozen_reference_checkpoint(name)
    sd = torch.load(buf, map_location="cpu")
    assert "min_val" not in "".join(sd.keys())             # the ranges are NOT in a checkpoint (quant_modules.py:133-134)
    got = Pack.from_state_dict(sd, num_heads=3) if name.startswith("deit") else Pack.from_state_dict({"model": sd})
    want = (export_deit if name.startswith("deit") else export_swin)(model)
    assert got.meta == want.meta
    assert set(got.arrays) == set(want.arrays)
    for k in want.arrays:
        assert got.arrays[k].dtype == want.arrays[k].dtype, k
        assert np.array_equal(got.arrays[k], want.arrays[k]), "array %s differs between checkpoint and live model" % k


def test_state_dict_saved_before_any_forward_is_refused():
    from ivit_b200.pack import from_state_dict
    m = refload.load()
    model = m.deit_tiny_patch16_224(pretrained=False).eval()
    with pytest.raises((ValueError, KeyError)):
        from_state_dict(model.state_dict(), num_heads=3)


def _reference_save_params():
    """The reference's own ``save_params`` (TVM_benchmark/convert_model.py:12-66), executed without importing the module
    (its top-level imports need tvm, which is not installed): the function's source is compiled on its own."""
    path = os.path.join(refload.REF, "TVM_benchmark", "convert_model.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "save_params"][0]
    ns = {"np": np, "os": os, "torch": torch, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["save_params"]


def test_export_tvm_params_matches_the_reference_converter(tmp_path):
    """params.npy written from the pack == params.npy written by the reference's converter from the same checkpoint
    (integer tensors exactly; cls / pos embeddings as integer * scale, which quantise back to the pack's integers)."""
    from ivit_b200.pack import export_tvm_params, from_state_dict
    model, buf = _frozen_reference_checkpoint("deit_tiny_patch16_224")
    sd = torch.load(buf, map_location="cpu")
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir()
    _reference_save_params()(sd, 12, str(ref_dir))
    ref = np.load(ref_dir / "params.npy", allow_pickle=True).item()
    pack = from_state_dict(sd, num_heads=3)
    params, q = export_tvm_params(pack, str(our_dir))
    ours = np.load(our_dir / "params.npy", allow_pickle=True).item()
    assert set(ours) == set(ref)
    for k, v in ref.items():
        if k in ("cls_token_weight", "pos_embed_weight"):
            continue
        assert ours[k].dtype == v.dtype and ours[k].shape == v.shape, k
        assert np.array_equal(ours[k], v), k
    s_pos = np.float32(pack["qact_pos.scale"][0])
    assert np.array_equal(np.clip(np.round(ours["pos_embed_weight"].reshape(-1) / s_pos), -32768, 32767),
                          pack["pos_embed_integer"].reshape(-1))
    # the scale chain: a few entries against the checkpoint's buffers (convert_model.py:80-148)
    assert q["qconfig_embed_conv"]["input_scale"] == float(sd["qact_input.act_scaling_factor"].reshape(-1)[0])
    assert np.allclose(q["block_3_qconfig_qkv"]["kernel_scale"], sd["blocks.3.attn.qkv.fc_scaling_factor"].numpy(), rtol=2e-7)
    assert q["block_3_qconfig_softmax"]["output_scale"] == float(sd["blocks.3.attn.int_softmax.act_scaling_factor"].reshape(-1)[0])
    assert q["block_11_qconfig_add2"]["output_scale"] == float(sd["blocks.11.qact4.act_scaling_factor"].reshape(-1)[0])
    assert np.array_equal(np.float32(q["qconfig_norm"]["output_scale"]), sd["norm.norm_scaling_factor"].numpy().reshape(-1))
