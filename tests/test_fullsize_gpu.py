"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle is too slow for a whole
bs=256 DeiT-B batch): images are independent, so the logits of the full batch must equal, bit for bit,
  * the logits of the same images run in sub-batches (other tile schedules, other M tails, other CUDA graphs),
  * an eager (un-graphed) run,
  * the CPU oracle on a slice of the batch.
Configs 2 and 3 of BASELINE.json: DeiT-small bs=128, DeiT-base bs=256 on one B200."""
import numpy as np
import pytest
import torch

import oracle.model as OM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,batch,sub", [("deit_small_patch16_224", 128, 48), ("deit_base_patch16_224", 256, 64)])
def test_full_batch_equals_subbatches_eager_run_and_oracle_slice(name, batch, sub):
    from ivit_b200.calib import build_synthetic
    from ivit_b200.engine import Engine
    from ivit_b200.pack import export_deit
    from ivit_b200.synth import synth_images
    pack = export_deit(build_synthetic(name))
    x = synth_images(batch, seed=77)
    xg = x.cuda()
    eng = Engine(pack, "cuda")
    full = eng(xg).clone()
    assert torch.isfinite(full).all()
    # sub-batches (the last one is ragged for 128 / 48)
    parts = [eng(xg[i:i + sub].contiguous()).clone() for i in range(0, batch, sub)]
    assert torch.equal(torch.cat(parts), full), "full batch differs from its sub-batches"
    # eager launches instead of the captured graph
    eager = Engine(pack, "cuda", use_cuda_graph=False)
    assert torch.equal(eager(xg), full), "graph replay differs from eager launches"
    # a slice against the CPU oracle (first, middle and last image)
    idx = [0, batch // 2, batch - 1]
    want = OM.deit_forward(pack, x[idx].numpy())
    assert np.array_equal(full[idx].cpu().numpy(), want), "engine logits differ from the oracle"
    # logits are not degenerate: images disagree with each other
    assert len(set(full.argmax(1).tolist())) > 1 or float((full[0] - full[1]).abs().max()) > 0


def test_swin_tiny_full_batch_equals_subbatches_eager_run_and_oracle_slice():
    """BASELINE.json config 4: Swin-tiny bs=128 through the fused SwinEngine (window attention on tcgen05, window glue as
    index math): full batch == sub-batches == eager launches == CPU oracle on a slice, and every attention launch took
    the tensor-core kernel."""
    from ivit_b200.calib import build_synthetic
    from ivit_b200.pack import export_swin
    from ivit_b200.swin_engine import SwinEngine
    from ivit_b200.synth import synth_images
    pack = export_swin(build_synthetic("swin_tiny_patch4_window7_224"))
    batch, sub = 128, 48
    x = synth_images(batch, seed=78)
    xg = x.cuda()
    eng = SwinEngine(pack, "cuda")
    full = eng(xg).clone()
    assert torch.isfinite(full).all()
    assert eng.attention_fallbacks == 0, "window attention fell back to the mma.sync kernel %d times" % eng.attention_fallbacks
    parts = [eng(xg[i:i + sub].contiguous()).clone() for i in range(0, batch, sub)]
    assert torch.equal(torch.cat(parts), full), "full batch differs from its sub-batches"
    eager = SwinEngine(pack, "cuda", use_cuda_graph=False)
    assert torch.equal(eager(xg), full), "graph replay differs from eager launches"
    idx = [0, batch // 2, batch - 1]
    want = OM.swin_forward(pack, x[idx].numpy())
    assert np.array_equal(full[idx].cpu().numpy(), want), "engine logits differ from the oracle"
    assert len(set(full.argmax(1).tolist())) > 1 or float((full[0] - full[1]).abs().max()) > 0
