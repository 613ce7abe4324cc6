"""freeze_model / unfreeze_model -- same contract as the reference's ``models/model_utils.py:5-40``:
recursively call ``fix()`` / ``unfix()`` on every QuantAct so that activation ranges stop / resume
updating.  (The reference walks ``dir(model)``; walking ``modules()`` reaches the same QuantActs,
also those inside ModuleList / Sequential containers.)"""
import torch.nn as nn

from .quantization_utils import QuantAct


def freeze_model(model: nn.Module):
    for m in model.modules():
        if type(m) in [QuantAct]:
            m.fix()


def unfreeze_model(model: nn.Module):
    for m in model.modules():
        if type(m) in [QuantAct]:
            m.unfix()
