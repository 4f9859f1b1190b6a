"""The import paths the reference scripts use (inference.py:8-12, ex_audioset.py:16-22) resolve to this
repository's modules when `dropin/` is on sys.path."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve_to_this_package():
    code = (
        "from models.mn.model import get_model as get_mobilenet\n"
        "from models.dymn.model import get_model as get_dymn\n"
        "from models.preprocess import AugmentMelSTFT\n"
        "from helpers.utils import NAME_TO_WIDTH, exp_warmup_linear_down, mixup\n"
        "import models.mn.model as m\n"
        "from models.ensemble import get_ensemble_model, EnsemblerModel\n"
        "from models.mn.model import get_ensemble_model as g2      # windowed_inference.py:8 spells it this way\n"
        "assert g2 is get_ensemble_model and 'efficientat_b200' in EnsemblerModel.__module__\n"
        "assert 'efficientat_b200' in get_mobilenet.__module__ and 'efficientat_b200' in AugmentMelSTFT.__module__\n"
        "assert NAME_TO_WIDTH('mn10_as') == 1.0 and NAME_TO_WIDTH('dymn20_as(2)') == 2.0 and NAME_TO_WIDTH('x') == 1.0\n"
        "net = get_mobilenet(width_mult=NAME_TO_WIDTH('mn04_as'), verbose=False)\n"
        "assert sum(p.numel() for p in net.parameters()) == 983599, sum(p.numel() for p in net.parameters())\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin") + os.pathsep + ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_param_counts_match_reference_table():
    """README.md:94-113 / SURVEY.md section 4: 4.88 M (mn10), 10.55 M (dymn10) parameters."""
    from efficientat_b200.models.dymn.model import get_model as dymn
    from efficientat_b200.models.mn.model import get_model as mn
    assert sum(p.numel() for p in mn(verbose=False).parameters()) == 4876831
    assert sum(p.numel() for p in dymn(verbose=False).parameters()) == 10548479


def test_window_plan_follows_the_reference_formula():
    """windowed_inference.py:95-97: n_windows = ceil((N - window) / hop) + 1, padded to n_windows * hop + window."""
    import numpy as np
    from efficientat_b200.windowed import window_plan
    for n, w, h in [(320000, 320000, 80000), (320000, 64000, 32000), (400001, 320000, 80000), (441000, 64000, 16000),
                    (63000, 64000, 32000)]:
        nw = int(np.ceil((n - w) / h)) + 1
        assert window_plan(n, w, h) == (nw, nw * h + w)
        assert nw * h + w >= n                                   # the reference's padding amount is never negative here
    assert window_plan(1000, 64000, 32000) == (1, 96000)         # documented difference: the reference yields 0 windows
    import pytest
    with pytest.raises(ValueError):
        window_plan(1000, 0, 10)


def test_ensemble_is_the_mean_of_member_logits_and_returns_it_twice():
    """models/ensemble.py:14-23 semantics of the native EnsemblerModel, with stand-in members (no GPU here)."""
    import torch
    from efficientat_b200.models.ensemble import EnsemblerModel

    class Member(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k = k

        def forward(self, x):
            return x.sum(dim=(1, 2, 3)).view(-1, 1) * self.k + torch.arange(4.0), None

    x = torch.randn(3, 1, 5, 7)
    ens = EnsemblerModel([Member(1.0), Member(2.0), Member(4.0)])
    a, b = ens(x)
    want = sum(m(x)[0] for m in ens.models) / 3
    assert a is b and torch.allclose(a, want, atol=1e-6)
    import pytest
    with pytest.raises(ValueError):
        EnsemblerModel([])(x)


def test_windowed_tagger_has_no_cpu_path():
    import pytest
    import torch
    from efficientat_b200.windowed import EATagger
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        EATagger(model_name="mn10_as", device="cuda")
    with pytest.raises(RuntimeError):
        EATagger(model_name="mn10_as", device="cpu")
