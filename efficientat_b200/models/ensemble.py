"""Drop-in for the reference's models/ensemble.py: an ensemble is the mean of its members' logits.

`get_ensemble_model(model_names)` builds every member through this package's `get_model` factories (MN / DyMN, with the
release checkpoints the names refer to, models/ensemble.py:26-35) and `EnsemblerModel.forward(x)` returns
`(mean_logits, mean_logits)` exactly like the reference (models/ensemble.py:14-23: the second element is NOT a feature
vector).  The members run back to back on the current stream; the average is one axpy per member on [B, classes]."""
import torch.nn as nn

from ..helpers.utils import NAME_TO_WIDTH
from .dymn.model import get_model as get_dymn
from .mn.model import get_model as get_mobilenet


class EnsemblerModel(nn.Module):
    def __init__(self, models):
        super().__init__()
        self.models = nn.ModuleList(models)

    def forward(self, x):
        if len(self.models) == 0:
            raise ValueError("EnsemblerModel needs at least one member")
        acc = None
        for m in self.models:
            out = m(x)[0]
            acc = out.clone() if acc is None else acc.add_(out)      # members may hand out views of their own buffers
        acc = acc.div_(len(self.models))
        return acc, acc


def get_ensemble_model(model_names):
    members = []
    for name in model_names:
        factory = get_dymn if name.startswith("dymn") else get_mobilenet
        members.append(factory(width_mult=NAME_TO_WIDTH(name), pretrained_name=name))
    return EnsemblerModel(members)
